"""Host logic of the scratch arena (palettenerf_b200/arena.py): a buffer is handed out again only when nothing references
it any more — Python variables, autograd contexts, saved tensors or views sharing its storage."""
import gc

import torch

from palettenerf_b200.arena import Arena


def _n(A, name, shape):
    cur, lst = A.pool[(name, torch.float32, "cpu")]
    assert cur == shape
    return len(lst)


def test_arena_reuses_only_unreferenced_buffers():
    A = Arena()
    a = A.get("x", (4,), torch.float32, "cpu")
    b = A.get("x", (4,), torch.float32, "cpu")
    assert a is not b and a.data_ptr() != b.data_ptr()              # both held -> two buffers
    pa = a.data_ptr()
    del a
    c = A.get("x", (4,), torch.float32, "cpu")
    assert c.data_ptr() == pa                                       # released -> reused
    view = c[:2]
    del c
    d = A.get("x", (4,), torch.float32, "cpu")
    assert d.data_ptr() not in (pa,)                                # a live view keeps the storage busy
    del view, d
    alias = A.get("x", (4,), torch.float32, "cpu").detach()         # the pattern used for autograd outputs
    e = A.get("x", (4,), torch.float32, "cpu")
    assert e.data_ptr() != alias.data_ptr()
    assert A.bytes() == sum(t.numel() * 4 for t in A.pool[("x", torch.float32, "cpu")][1])


def test_arena_respects_autograd_lifetimes():
    A = Arena()

    class Saved(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            y = A.get("y", (3,), torch.float32, "cpu").detach()
            y.copy_(x * 2)
            ctx.save_for_backward(y)
            return y

        @staticmethod
        def backward(ctx, g):
            return g * 2

    class Kept(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            ctx.keep = (A.get("k", (3,), torch.float32, "cpu"),)
            return x * 2

        @staticmethod
        def backward(ctx, g):
            return g * 2

    x = torch.ones(3, requires_grad=True)
    for _ in range(3):                       # steady state: one buffer per name, reused every step
        out = Saved.apply(x) + Kept.apply(x)
        out.sum().backward()
        del out
        gc.collect()
    assert _n(A, "y", (3,)) == 1 and _n(A, "k", (3,)) == 1
    o1 = Saved.apply(x)                      # graph alive -> a second forward must not clobber the first one's buffers
    o2 = Saved.apply(x)
    assert o1.data_ptr() != o2.data_ptr() and _n(A, "y", (3,)) == 2


def test_arena_drops_buffers_of_a_previous_shape():
    """the sample capacity follows mean_count, which changes at every density-grid refresh: buffers of the old capacity
    must not pile up (round-1 advisor finding)"""
    A = Arena()
    for cap in (1000, 1200, 900, 1500):
        t = A.get("xbuf", (cap, 8), torch.float32, "cpu")
        assert t.shape == (cap, 8)
        del t
        assert A.bytes() == cap * 8 * 4                               # only the live shape is pooled
    a = A.get("xbuf", (1500, 8), torch.float32, "cpu")
    b = A.get("xbuf", (700, 8), torch.float32, "cpu")               # another shape while `a` is still held
    assert a.shape == (1500, 8) and b.shape == (700, 8) and A.bytes() == 700 * 8 * 4
