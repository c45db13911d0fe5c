"""Import-time stand-in (never executed on the hot path)."""


def scatter_min(*a, **k):
    raise NotImplementedError
