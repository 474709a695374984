"""oqupy_b200 -- B200-native TEMPO / PT-TEMPO engine behind OQuPy's backend API.

Only the hot path lives here: device-resident MPS chain,
influence-MPO x MPS zip-up, eps-truncated SVD sweeps, process-tensor export and
the compute_dynamics loop.  Everything reaches the GPU through the C-ABI library
``liboqupy_b200.so`` (include/oqupy_b200.h); there is no CPU fallback.
"""
from .backends import (BaseTempoBackend, MeanFieldTempoBackend, PtTempoBackend,
                       TempoBackend)
from .tebd import PtTebdBackend
from .process_tensor import (DeviceProcessTensor, dynamics_device,
                             dynamics_with_field_device, gradient_device,
                             import_process_tensor)
from ._lib import B200Error, CudaOps, default_ops, load_library
from .batch import BatchedTempoBackend

__all__ = ["BatchedTempoBackend", "BaseTempoBackend", "MeanFieldTempoBackend", "PtTempoBackend", "TempoBackend",
           "PtTebdBackend",
           "DeviceProcessTensor", "dynamics_device", "dynamics_with_field_device", "gradient_device",
           "import_process_tensor", "B200Error", "CudaOps",
           "default_ops", "load_library"]
__version__ = "0.1.0"
