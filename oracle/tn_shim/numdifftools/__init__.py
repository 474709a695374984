"""TEST INFRASTRUCTURE ONLY: import stub (system.py:25)."""


class Jacobian:  # pylint: disable=too-few-public-methods
    def __init__(self, *a, **k):
        raise ImportError("numdifftools is stubbed in the oracle environment")
