"""No-op matplotlib (test infrastructure): the reference's driver and post-processing scripts import
it for plots that are outside the scan/scoring path; it is not installed in this image."""


class _Nothing:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __iter__(self):
        return iter(())

    def __getitem__(self, i):
        return self


_nothing = _Nothing()


def use(*a, **k):
    return None


def __getattr__(name):
    return _nothing
