"""TEST INFRASTRUCTURE ONLY (type-annotation stub, node_array.py:21)."""


class BaseBackend:  # pylint: disable=too-few-public-methods
    pass
