"""stand-in for cvxopt (convex-hull palette extraction only)"""


def __getattr__(name):
    def _unavailable(*args, **kwargs):
        raise NotImplementedError(f"cvxopt.{name}: cvxopt is not installed (compat stand-in)")
    return _unavailable
