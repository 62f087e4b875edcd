"""stand-in for trimesh (mesh export / debug visualisation only)"""


class Trimesh:
    def __init__(self, vertices=None, faces=None, **kwargs):
        self.vertices, self.faces = vertices, faces

    def export(self, path, **kwargs):
        print(f"[compat.trimesh] trimesh is not installed in this image: {path} not written")


def __getattr__(name):
    def _unavailable(*args, **kwargs):
        raise NotImplementedError(f"trimesh.{name}: trimesh is not installed (compat stand-in)")
    return _unavailable
