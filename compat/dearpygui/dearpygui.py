def __getattr__(name):
    def _unavailable(*args, **kwargs):
        raise NotImplementedError(f"dearpygui.{name}: no GUI in this image (compat stand-in)")
    return _unavailable
