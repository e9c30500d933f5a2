"""Test-harness stub: reference focusr.py:11 does `from matplotlib import colors`."""


class colors:  # noqa: N801
    pass
