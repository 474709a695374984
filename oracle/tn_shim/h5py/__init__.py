"""TEST INFRASTRUCTURE ONLY: import stub (h5py is not installable here).
Only FileProcessTensor (process_tensor.py:433-881, out of scope) touches it."""


def __getattr__(name):
    raise ImportError("h5py is stubbed in the oracle environment")
