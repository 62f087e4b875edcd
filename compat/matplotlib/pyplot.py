def _unavailable(*args, **kwargs):
    raise NotImplementedError("matplotlib is not installed in this image (compat stand-in)")


def __getattr__(name):
    return _unavailable
