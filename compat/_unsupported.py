"""names the reference's scripts import but that lie outside the B200 hot path: importable, unusable"""


def unsupported(name, why="outside the B200 hot path (SURVEY.md 8): use the reference implementation"):
    class _Unsupported:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name}: {why}")
    _Unsupported.__name__ = _Unsupported.__qualname__ = name
    return _Unsupported
