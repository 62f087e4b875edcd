import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def load_ref(name):
    """import a reference extension built by oracle/build_ref.py (oracle/_ref/_ref_<name>.so) or return None"""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    path = os.path.join(ROOT, "oracle", "_ref", f"_ref_{name}.so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def scene():
    """lego-shaped synthetic scene on the CPU: density grid, bitfield, aabb"""
    import torch
    from palettenerf_b200 import synthetic as S
    grid = S.density_grid()
    thresh = min(grid.clamp(min=0).mean().item(), S.LEGO["density_thresh"])
    bitfield = S.packbits_np(grid, thresh)
    aabb = torch.tensor([-2, -2, -2, 2, 2, 2], dtype=torch.float32)
    return dict(grid=grid, thresh=thresh, bitfield=bitfield, aabb=aabb, **S.LEGO)
