"""``gstools_cython.variogram`` stand-in (out of scope; variogram/variogram.py:14-17)."""


def _unavailable(*args, **kwargs):
    raise NotImplementedError("variogram estimation is outside the summator hot path")


directional = unstructured = structured = ma_structured = _unavailable
