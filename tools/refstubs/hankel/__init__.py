"""``hankel`` stand-in: models without analytic spectral density are unavailable offline."""


class SymmetricFourierTransform:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs

    def transform(self, *args, **kwargs):
        raise NotImplementedError("hankel is not installed in this environment")
