"""TEST INFRASTRUCTURE ONLY: import stub."""


class Axes:  # pylint: disable=too-few-public-methods
    pass
