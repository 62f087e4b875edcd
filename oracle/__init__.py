"""oracle — CPU restatement of the reference's algorithms for the hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
the product (palettenerf_b200/) never does. See oracle/raymarch_oracle.c for the parity-pinning statement: the
reference ships no golden vectors for this path, so the oracle is pinned against the reference's own kernels
(oracle/_ref, built from /root/reference by oracle/build_ref.py) on the GPU box (tests/test_golden.py and the *_gpu tests)
and against the reference's own Python run on those kernels (tests/golden/ref_palette.npz, tests/test_golden_palette.py).
"""
from .cpu_oracle import *  # noqa: F401,F403
