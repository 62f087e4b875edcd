"""FusedAdam (the optimizer of ref palette/utils.py:719-724) vs torch.optim.Adam on the same
parameters and gradients: fp32, |a - b| <= 2e-6 |b| + 2e-8 on the parameters after several steps (2e-8 = 2e-6 of one
lr-sized update; torch's own fused and foreach Adam differ by the same order), exact agreement on skipped steps and step
counters, state_dict interchange."""
import pytest
import torch


def _close(a, b, rel=2e-6, abs_=2e-8):
    return bool(((a - b).abs() <= rel * b.abs() + abs_).all())


def test_fused_adam_needs_cuda():
    from palettenerf_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedAdam([p], lr=1e-2).step()
    with pytest.raises(RuntimeError):
        FusedAdam([p], amsgrad=True)


def _params(dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1 << 16, 2), (64, 32), (16, 64), (3,), (4, 3), (1000003,), (5,)]     # unaligned / odd sizes included
    return [torch.randn(*s, generator=g).to(dev) for s in shapes]


@pytest.mark.gpu
@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_fused_adam_matches_torch_adam(cuda, wd):
    from palettenerf_b200.optim import FusedAdam
    base = _params(cuda)
    pa = [torch.nn.Parameter(t.clone()) for t in base]
    pb = [torch.nn.Parameter(t.clone()) for t in base]
    kw = dict(lr=1e-2, betas=(0.9, 0.99), eps=1e-15, weight_decay=wd)          # the reference's Adam
    ref = torch.optim.Adam(pb, **kw)
    opt = FusedAdam(pa, **kw)
    g = torch.Generator().manual_seed(1)
    for it in range(6):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).to(cuda) * (10.0 ** (it - 3))
            if it == 2 and a.dim() == 2:
                gr[0] = 0.0                                                     # untouched entries: g = 0
            a.grad, b.grad = gr.clone(), gr.clone()
        pa[-1].grad = None if it == 4 else pa[-1].grad                          # a parameter without a gradient is skipped
        pb[-1].grad = None if it == 4 else pb[-1].grad
        opt.step(); ref.step()
    for a, b in zip(pa, pb):
        assert _close(a, b), (a - b).abs().max().item()
    for a, b in zip(pa, pb):
        sa, sb = opt.state[a], ref.state[b]
        assert float(sa["step"]) == float(sb["step"])
        # moments: relative to the largest entry (an exp_avg entry can cancel to ~0 from O(10) gradients)
        for k in ("exp_avg", "exp_avg_sq"):
            assert _close(sa[k], sb[k], 2e-6, 2e-6 * sb[k].abs().max().item()), k


@pytest.mark.gpu
def test_fused_adam_with_grad_scaler_skips_on_inf_and_interchanges_state(cuda):
    from palettenerf_b200.optim import FusedAdam
    base = _params(cuda, seed=2)[:4]
    pa = [torch.nn.Parameter(t.clone()) for t in base]
    pb = [torch.nn.Parameter(t.clone()) for t in base]
    kw = dict(lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    opt, ref = FusedAdam(pa, **kw), torch.optim.Adam(pb, fused=True, capturable=True, **kw)
    sa, sb = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    g = torch.Generator().manual_seed(3)
    for it in range(5):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).to(cuda) * 1024.0 * (0.5 if it > 2 else 1.0)   # scale halves after the skip
            if it == 2 and a is pa[0]:
                gr.view(-1)[7] = float("inf")
            a.grad, b.grad = gr.clone(), gr.clone()
        before = [a.detach().clone() for a in pa]
        sa.scale(torch.ones((), device=cuda)); sb.scale(torch.ones((), device=cuda))   # lazy-initialises the scalers
        sa.step(opt); sa.update()
        sb.step(ref); sb.update()
        if it == 2:
            assert all(torch.equal(x, y) for x, y in zip(before, pa))          # skipped on the device
            assert float(opt.state[pa[0]]["step"]) == 2.0
    assert sa.get_scale() == sb.get_scale() == 512.0
    for a, b in zip(pa, pb):
        assert _close(a, b), (a - b).abs().max().item()
        assert float(opt.state[a]["step"]) == float(ref.state[b]["step"]) == 4.0
    # state interchange: torch -> FusedAdam and back
    opt2 = FusedAdam(pa, **kw)
    opt2.load_state_dict(ref.state_dict())
    ref2 = torch.optim.Adam(pb, fused=True, capturable=True, **kw)
    ref2.load_state_dict(opt.state_dict())
    for a, b in zip(pa, pb):
        gr = torch.randn(a.shape, generator=g).to(cuda)
        a.grad, b.grad = gr.clone(), gr.clone()
    opt2.step(); ref2.step()
    for a, b in zip(pa, pb):
        assert _close(a, b, 4e-6, 4e-8), (a - b).abs().max().item()


@pytest.mark.gpu
def test_fused_adam_in_a_cuda_graph_with_device_lr(cuda):
    from palettenerf_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.randn(4097, device=cuda))
    q = torch.nn.Parameter(p.detach().clone())
    lr = torch.tensor(1e-2, device=cuda)
    opt = FusedAdam([p], lr=lr, betas=(0.9, 0.99), eps=1e-15)
    ref = torch.optim.Adam([q], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    grad = torch.randn(4097, device=cuda)
    p.grad = grad.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        opt.step()                                # state allocation outside the capture
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt.step()
    q.grad = grad.clone(); ref.step()            # the eager step; the captured one only runs on replay
    for k in range(3):
        lr.fill_(1e-2 * 0.5 ** (k + 1))           # the schedule reaches the captured kernel through device memory
        graph.replay()
        for gq in ref.param_groups:
            gq["lr"] = 1e-2 * 0.5 ** (k + 1)
        ref.step()
    assert float(opt.state[p]["step"]) == 4.0
    assert _close(p, q, 4e-6, 4e-8), (p - q).abs().max().item()


@pytest.mark.gpu
def test_fused_adam_bumps_tensor_versions(cuda):
    """the kernel writes parameters through raw pointers; version-keyed caches (the renderer's fp16 table copies) rely on
    `_version` changing like it does for torch's in-place optimizers"""
    from palettenerf_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.randn(1000, device=cuda))
    opt = FusedAdam([p], lr=1e-2)
    p.grad = torch.randn(1000, device=cuda)
    v0 = p._version
    opt.step()
    assert p._version > v0


@pytest.mark.gpu
def test_grad_scaler_with_fused_found_inf_check_behaves_like_torch(cuda):
    """optim.GradScaler: the non-finite check in front of a FusedAdam step is pnerf_found_inf (one streaming pass) instead of
    torch's multi-tensor check-and-unscale; same skips, same scale schedule, same parameters as torch's scaler"""
    from palettenerf_b200.optim import FusedAdam, GradScaler
    g = torch.Generator().manual_seed(11)
    shapes = [(1 << 20) + 3, 1, 7, 64 * 64, 1023] + [5 + i for i in range(36)]       # odd sizes, > 32 tensors, one large
    base = [torch.randn(s, generator=g).to(cuda) for s in shapes]
    base[4] = torch.randn(1024, generator=g).to(cuda)[1:]                               # a 4-byte-aligned (not 16) gradient
    pa = [torch.nn.Parameter(t.clone()) for t in base]
    pb = [torch.nn.Parameter(t.clone()) for t in base]
    kw = dict(lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    opt, ref = FusedAdam(pa, **kw), FusedAdam(pb, **kw)
    sa, sb = GradScaler("cuda", init_scale=256.0, growth_interval=3), torch.amp.GradScaler("cuda", init_scale=256.0, growth_interval=3)
    poison = {1: (0, 12345, float("inf")), 3: (40, 2, float("nan")), 5: (4, 1022, float("-inf")), 6: (0, (1 << 20) + 2, float("nan"))}
    for it in range(9):
        for k, (a, b) in enumerate(zip(pa, pb)):
            gr = torch.randn(a.shape, generator=g).to(cuda)
            if gr.data_ptr() % 16 == 0 and k == 4:
                gr = torch.randn(a.numel() + 1, generator=g).to(cuda)[1:]
            if it in poison and poison[it][0] == k:
                gr.view(-1)[poison[it][1]] = poison[it][2]
            a.grad, b.grad = gr.clone() if k != 4 else gr, gr.clone()
        before = [a.detach().clone() for a in pa]
        sa.scale(torch.ones((), device=cuda)); sb.scale(torch.ones((), device=cuda))
        sa.step(opt); sa.update()
        sb.step(ref); sb.update()
        if it in poison:
            assert all(torch.equal(x, y) for x, y in zip(before, pa)), it
        else:
            assert not torch.equal(before[0], pa[0])
        assert sa.get_scale() == sb.get_scale(), (it, sa.get_scale(), sb.get_scale())
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)


def test_grad_scaler_subclass_falls_back_to_torch_for_other_optimizers():
    """optim.GradScaler only replaces the non-finite check in front of a FusedAdam step; any other optimizer (here on the CPU)
    goes through torch's own path unchanged"""
    from palettenerf_b200.optim import GradScaler
    p = torch.nn.Parameter(torch.ones(5))
    opt = torch.optim.SGD([p], lr=0.1)
    sc = GradScaler("cpu", init_scale=4.0, growth_interval=2)
    for it, bad in enumerate((False, True, False, False)):
        opt.zero_grad()
        loss = (p * (float("inf") if bad else 1.0)).sum()
        sc.scale(loss).backward()
        before = p.detach().clone()
        sc.step(opt)
        sc.update()
        assert torch.equal(before, p.detach()) == bad, it
    assert sc.get_scale() == 4.0          # halved by the skipped step, doubled again after two good ones
