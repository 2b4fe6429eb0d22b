"""``meshio`` stand-in (field/tools.py:14)."""


class Mesh:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("meshio is not installed in this environment")
