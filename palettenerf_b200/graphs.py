"""CUDA-graph capture of a whole training step (forward + loss + backward + optimizer) of the fused palette path.

The reference's training step is launch-bound: ~250 kernel launches of a few microseconds each per 4096-ray batch plus
a D2H synchronisation inside march_rays_train (raymarching/raymarching.py:224). The fused training path of this
repository has data-independent shapes and no host synchronisation (static-capacity march, device-side sample count),
so the step is captured ONCE into a CUDA graph and replayed: one graph launch per step, the GPU never waits for Python.

Requirements on `step_fn` (checked by use, not statically): no .item()/.cpu() inside, optimizer built with
`capturable=True` (or `fused=True`), gradients created inside the capture (`zero_grad(set_to_none=True)` first), inputs
read from the static tensors returned by `static_inputs` (copy new data into them before `replay()`).
"""
import torch


class GraphedStep:
    def __init__(self, step_fn, warmup=3, pool=None):
        """step_fn() runs one full training step on the CURRENT stream and returns a tensor (e.g. the loss) or None.
        It is executed `warmup` times eagerly on a side stream, then captured."""
        self.step_fn = step_fn
        self.graph = torch.cuda.CUDAGraph()
        self.out = None
        # Warm-up and capture run on ONE side stream: autograd's AccumulateGrad nodes remember the stream they were
        # created on, and a node that lives on the legacy default stream would make the capture depend on it
        # (cudaErrorStreamCaptureImplicit). For the same reason the model must not have been stepped on the default
        # stream before it is handed to GraphedStep.
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                step_fn()
        torch.cuda.current_stream().wait_stream(self.stream)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph, pool=pool, stream=self.stream):
            self.out = step_fn()

    def replay(self):
        self.graph.replay()
        return self.out


def make_palette_train_step(model, optimizer, scaler, rays_o, rays_d, loss_fn, render_kwargs=None, bucket=None):
    """-> step_fn for GraphedStep: palette-stage step on the static tensors rays_o / rays_d ([1, N, 3]).
    loss_fn(out) -> scalar loss from the render dict. bucket: optional distributed.GradBucket (one all-reduce per step)."""
    kw = dict(staged=False, bg_color=1, perturb=True, force_all_rays=True, dt_gamma=0.0, max_steps=1024)
    kw.update(render_kwargs or {})
    if bucket is not None:
        # the fused backward places the hash-table gradients in the bucket itself and starts their all-reduce early
        object.__setattr__(model, "_grad_bucket", bucket)

    def step_fn():
        optimizer.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.render(rays_o, rays_d, **kw)
            loss = loss_fn(out)
        scaler.scale(loss).backward()
        if bucket is not None:
            bucket.all_reduce(average=True)
        scaler.step(optimizer)
        scaler.update()
        return loss.detach()
    return step_fn


def make_nerf_train_step(model, optimizer, scaler, rays_o, rays_d, gt_rgb, lambda_sparse=0.05, render_kwargs=None, bucket=None):
    """-> step_fn for GraphedStep: stage-1 step (ref: Trainer.train_step, nerf/utils.py:485-560 with the MSE criterion) on the
    static tensors rays_o / rays_d / gt_rgb ([1, N, 3]). bucket: optional distributed.GradBucket (one all-reduce per step; the
    fused backward places the density table's gradient in it and starts that region's all-reduce early)."""
    kw = dict(staged=False, bg_color=1, perturb=True, force_all_rays=True, dt_gamma=0.0, max_steps=1024)
    kw.update(render_kwargs or {})
    if bucket is not None:
        object.__setattr__(model, "_grad_bucket", bucket)

    def step_fn():
        optimizer.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.render(rays_o, rays_d, rays_gt=gt_rgb, **kw)
            loss = (((out["image"] - gt_rgb) ** 2).mean(-1) + lambda_sparse * out["rgb_norm"]).mean()
        scaler.scale(loss).backward()
        if bucket is not None:
            bucket.all_reduce(average=True)
        scaler.step(optimizer)
        scaler.update()
        return loss.detach()
    return step_fn
