"""stand-in for PyMCubes (mesh extraction only)"""


def marching_cubes(*args, **kwargs):
    raise NotImplementedError("PyMCubes is not installed in this image (compat stand-in)")
