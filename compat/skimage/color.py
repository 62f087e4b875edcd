def __getattr__(name):
    def _unavailable(*args, **kwargs):
        raise NotImplementedError(f"skimage.color.{name}: scikit-image is not installed (compat stand-in)")
    return _unavailable
