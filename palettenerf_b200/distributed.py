"""Multi-GPU plumbing of the hot path: one process per GPU, `torch.distributed` (NCCL over NVLink on
the box, gloo in the CPU tests). Rays are independent, so

  * inference shards the rays (`ray_shard`) — contiguous bands, or interleaved tiles when empty-space rays make
    bands unbalanced. `ShardedView` renders ONE view on all ranks: the persistent kernel of every rank stores its finished
    rays straight into the owner's image in symmetric (peer-mapped) memory — compute and gather in one kernel;
    `gather_maps` is the library-collective alternative (all-gather of the per-ray maps, gloo-testable);
  * training is ray-batch data parallel: `GradBucket` packs the gradients of the tensors that actually received one
    (the sigma grid / sigma net get none in the palette stage, palette/network.py:168, palette/renderer.py:335) plus
    the GradScaler found-inf flag into ONE flat fp32 buffer and all-reduces it once per step;
  * the stochastic density-grid refresh (nerf/renderer.py:502-519) would make ranks diverge: `cell_shard` splits the
    (cascade, cell) space between ranks and `merge_density` combines the partial grids with an all-reduce(max).

The reference is single-GPU; none of this has a counterpart there.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def ray_shard(n_rays, world_size, rank, tile=0):
    """indices of the rays this rank renders.
    tile == 0: contiguous band [lo, hi) -> returns a slice; tile > 0: round-robin blocks of `tile` rays -> LongTensor."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    if tile <= 0:
        per = (n_rays + world_size - 1) // world_size
        lo = min(rank * per, n_rays)
        return slice(lo, min(lo + per, n_rays))
    blocks = torch.arange((n_rays + tile - 1) // tile)
    mine = blocks[blocks % world_size == rank]
    idx = (mine[:, None] * tile + torch.arange(tile)[None, :]).reshape(-1)
    return idx[idx < n_rays]


def gather_maps(local, n_rays, shard, group=None, tile=0):
    """all-gather a per-ray tensor [n_local, ...] rendered for `shard` (= ray_shard(n_rays, ws, rank, tile)) into
    [n_rays, ...] on every rank: ONE all_gather of equal-sized (padded) shards — every rank receives n_rays rows in total,
    not world x n_rays as a sum of zero-padded images would cost"""
    ws, rank = world()
    if ws == 1:
        full = torch.zeros((n_rays,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        full[shard] = local
        return full
    sizes = []
    for r in range(ws):
        sh = ray_shard(n_rays, ws, r, tile)
        sizes.append(sh.stop - sh.start if isinstance(sh, slice) else sh.numel())
    cap = max(sizes)
    padded = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(parts, padded, group=group)
    full = torch.zeros((n_rays,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(ws):
        sh = ray_shard(n_rays, ws, r, tile)
        full[sh if isinstance(sh, slice) else sh.to(local.device)] = parts[r][: sizes[r]]
    return full


class ShardedView:
    """ONE view rendered by all ranks of a box (north_star: "rendering is split by ray/image tile across the 8 GPUs").

    The rays of the view are dealt out in interleaved tiles of `tile` consecutive pixels (empty-space rays cost nothing, so
    contiguous bands would be unbalanced); every rank renders its shard with the persistent tensor-core renderer, whose
    `retire` step stores each finished ray STRAIGHT INTO THE OWNER'S output maps: they live in symmetric memory (every rank
    maps the owner's buffer, torch `_symmetric_memory`), so the stores travel over NVLink and the kernel itself is the
    gather — no all-gather / all-reduce, no staging copy. Two cross-GPU barriers of the symmetric-memory handle bracket the
    kernel (maps zeroed before anyone writes / every shard delivered); the per-ray epilogue (background mix, depth
    normalisation) runs on the owner. A ray's result does not depend on which rays share its launch, so the assembled
    image equals the single-GPU image bit for bit (tests/test_peer_gpu.py)."""

    def __init__(self, model, n_rays, gui_mode=False, tile=256, owner=0, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import fused
        group = group if group is not None else dist.group.WORLD
        self.model, self.n_rays, self.gui_mode, self.owner = model, n_rays, gui_mode, owner
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        dev = model.encoder.embeddings.device
        self.layout, total = fused.accumulator_layout(n_rays, model.num_basis, model.opt.clip_dim, gui_mode)
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:   # noqa: BLE001
            pass
        self.buf = symm.empty(total, dtype=torch.float32, device=dev)
        self.handle = symm.rendezvous(self.buf, group)
        owner_buf = self.buf if self.rank == owner else self.handle.get_buffer(owner, (total,), torch.float32)
        self.maps = {k: owner_buf[o:o + int(torch.Size(shp).numel())].view(*shp) for k, (o, shp) in self.layout.items()}
        self.shard = ray_shard(n_rays, self.world, self.rank, tile)
        self.shard_dev = self.shard.to(dev)
        self.out_index = self.shard_dev.to(torch.int32).contiguous()

    @torch.no_grad()
    def render(self, rays_o, rays_d, bg_color=1, perturb=False, dt_gamma=0.0, max_steps=1024, T_thresh=1e-4):
        """rays_o, rays_d: [n_rays, 3] of the WHOLE view on every rank (generated from the pose: 64 bytes of input).
        Returns run_cuda's inference dict on the owner rank, None elsewhere."""
        from . import fused, raymarching
        from .nerf.renderer import render_tail
        m = self.model
        o, d = rays_o.view(-1, 3)[self.shard_dev].contiguous(), rays_d.view(-1, 3)[self.shard_dev].contiguous()
        nears, fars = raymarching.near_far_from_aabb(o, d, m.aabb_infer, m.min_near)
        if self.rank == self.owner:
            self.buf.zero_()
        self.handle.barrier(channel=0)                      # the owner's maps are zero before any rank stores into them
        acc = fused.render(m, o, d, nears, fars, perturb, dt_gamma, max_steps, T_thresh, self.gui_mode, kernel="tc",
                           out=self.maps, out_index=self.out_index)
        self.handle.barrier(channel=1)                      # every shard has been delivered
        self.queue = acc["_queue"]
        if self.rank != self.owner:
            return None
        nears_all, fars_all = raymarching.near_far_from_aabb(rays_o.view(-1, 3), rays_d.view(-1, 3), m.aabb_infer, m.min_near)
        mp = self.maps
        depth_n, image, direct = render_tail(mp["depth"], nears_all, fars_all, mp["image"], mp["weights_sum"], bg_color,
                                             None if self.gui_mode else mp["direct_rgb"], 0)
        out = {"depth": depth_n, "depth_origin": mp["depth"], "image": image, "weights_sum": mp["weights_sum"],
               "clip_feat": mp["clip_feat"]}
        if not self.gui_mode:
            out.update(direct_rgb=direct, view_dep_rgb=mp["view_dep_rgb"], basis_rgb=mp["basis_rgb"],
                       unscaled_basis_rgb=mp["unscaled_basis_rgb"], basis_acc=mp["basis_acc"])
        return out


class PeerMemory:
    """the flat bucket as SYMMETRIC memory: every rank of the box maps every other rank's bucket (torch's
    `_symmetric_memory`: CUDA VMM / IPC handles exchanged once at rendezvous), so the all-reduce is our own kernel over
    NVLink loads and stores (csrc/peer.cu::pnerf_peer_allreduce) between two cross-GPU barriers of the same handle."""

    def __init__(self, n_floats, device, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("peer all-reduce: at most 8 ranks (one NVSwitch box)")
        pad = 4 * self.world
        self.n = (n_floats + pad - 1) // pad * pad
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:   # noqa: BLE001  (newer torch enables every group implicitly)
            pass
        self.buf = symm.empty(self.n, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        import ctypes
        import os
        self._ptrs = (ctypes.c_uint64 * self.world)(*[int(p) for p in self.handle.buffer_ptrs])
        self._range_ptrs = {0: self._ptrs}
        # NVLS: the bucket's multicast mapping (0 when the fabric / driver has no multicast support) -> in-switch reduction.
        # Used from 4 ranks up: per GPU it moves 1/world of the bucket instead of (world-1)/world; at 2 ranks the traffic
        # is the same and the multimem round trip is slower (measured: 0.771 vs 0.721 ms per DP step, profiles/README.md).
        want_mc = os.environ.get("PNERF_PEER_MULTICAST", "auto")
        use_mc = want_mc == "1" or (want_mc == "auto" and self.world >= 4)
        self.mc_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0) if use_mc else 0

    def all_reduce_(self, average, lo=0, hi=None, channels=(0, 1)):
        """all-reduce of bucket[lo:hi] (both multiples of 4 * world) on the CURRENT stream. All calls of a process must be
        stream-ordered with respect to each other (see GradBucket.early)"""
        import ctypes
        from . import _lib as L
        hi = self.n if hi is None else hi
        pad = 4 * self.world
        if lo % pad or hi % pad or not (0 <= lo < hi <= self.n):
            raise ValueError(f"peer all-reduce range [{lo}, {hi}) must be a multiple of {pad} inside the bucket")
        scale = (1.0 / self.world) if average else 1.0
        self.handle.barrier(channel=channels[0])            # every rank's gradients are in its bucket
        if self.mc_ptr:
            L.call("pnerf_peer_allreduce_mc", self.mc_ptr + 4 * lo, self.world, self.rank, hi - lo, scale, L.stream())
        else:
            ptrs = self._range_ptrs.get(lo)
            if ptrs is None:
                ptrs = self._range_ptrs[lo] = (ctypes.c_uint64 * self.world)(*[int(p) + 4 * lo for p in self.handle.buffer_ptrs])
            L.call("pnerf_peer_allreduce", ctypes.addressof(ptrs), self.world, self.rank, hi - lo, scale, L.stream())
        self.handle.barrier(channel=channels[1])            # every slice has been delivered to every rank


class GradBucket:
    """ONE flat fp32 bucket for all trainable gradients (+ the found-inf flag of the loss scaler), all-reduced once per step.
    peer=True: the bucket lives in symmetric memory and the all-reduce is `pnerf_peer_allreduce[_mc]` over NVLink peer
    memory (CUDA, one box, <= 8 ranks); otherwise (and on CPU / gloo) `dist.all_reduce`.

    Layout: the LARGE tensors (>= 2^20 elements: the hash tables, 99.9 % of the bytes) first, padded to the collective's
    granularity, then the small ones and the flag. The fused backward (palettenerf_b200/fused_train.py) asks `slot(p)` for
    the place of a table's gradient, scatters straight into it — no pack copy — and calls `early(params)`: the large region
    is then all-reduced on a SIDE STREAM while the main stream still computes the MLP weight gradients; `all_reduce()`
    afterwards packs and reduces the small region and joins the side stream."""
    BIG = 1 << 20

    def __init__(self, params, peer=False):
        self.params = [p for p in params if p.requires_grad]
        self.flat = None
        self._sizes, self._views, self._off = None, [], {}
        self.peer, self._pm = bool(peer), None
        self._big_end = 0
        self._early_event, self._side = None, None
        self.early_count = 0                                # times the large region went out early (tests / bench report it)

    def _live(self):
        live = [p for p in self.params if p.grad is not None]
        return sorted(live, key=lambda p: 0 if p.numel() >= self.BIG else 1)          # stable: big tensors first

    # ---- direct placement + early all-reduce of the big region (peer path only) ----
    def slot(self, p):
        """fresh view of the bucket where p's gradient lives (None until the first all_reduce has laid the bucket out, or
        when p is not one of the large tensors of the peer path)"""
        off = self._off.get(id(p))
        if off is None or self._pm is None or off >= self._big_end:
            return None
        return self.flat[off:off + p.numel()].view_as(p)

    def early(self, params, average=True):
        """the gradients of `params` have been written into their slots: if they are the whole large region, all-reduce it
        now on the side stream (returns True), else do nothing"""
        ws, _ = world()
        if self._pm is None or ws <= 1 or self._big_end == 0:
            return False
        if sum(p.numel() for p in params) != self._big_numel or any(self._off.get(id(p), self._big_end) >= self._big_end for p in params):
            return False
        # EVERY cross-GPU barrier of a step is issued on the side stream, in the same order on every rank (this region, then
        # the small one in all_reduce()): two spin-waiting barriers left unordered on one rank — one per stream — can
        # deadlock against a peer whose runtime happens to schedule them the other way round (seen under CUDA graphs)
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            self._pm.all_reduce_(average, 0, self._big_end)
        self._early_event = True
        self.early_count += 1
        return True

    def all_reduce(self, found_inf=None, average=True, group=None):
        """sums (or averages) gradients across ranks; returns the global found-inf flag (max over ranks).
        Pack = ONE multi-tensor copy of the gradients that are not already in place; after the collective every `p.grad` IS
        its slice of the bucket (re-pointed, not copied back). The bucket is owned by the gradients until the next backward."""
        ws, _ = world()
        live = self._live()
        sizes = [p.grad.numel() for p in live]
        dev = live[0].grad.device if live else torch.device("cpu")
        pad = 4 * ws
        big = sum(k for p, k in zip(live, sizes) if p.numel() >= self.BIG)
        big_end = (big + pad - 1) // pad * pad if (self.peer and ws > 1 and dev.type == "cuda") else big
        n = big_end + (sum(sizes) - big) + 1
        if self.flat is None or self._sizes != sizes or self.flat.device != dev:
            self._pm = None
            if self.peer and ws > 1 and dev.type == "cuda":
                try:
                    self._pm = PeerMemory(n, dev, group)
                except Exception as e:   # noqa: BLE001  (no symmetric memory on this box / build: the NCCL path still works)
                    import warnings
                    warnings.warn(f"peer all-reduce unavailable ({type(e).__name__}: {e}); using dist.all_reduce")
            if self._pm is not None:
                self.flat = self._pm.buf[:n]
            else:
                self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
            self._sizes = sizes
            self._views, self._off, off = [], {}, 0
            for p, k in zip(live, sizes):
                if off == big and big_end != big:
                    off = big_end                                  # (padding between the large and the small region)
                self._views.append(self.flat[off:off + k])
                self._off[id(p)] = off
                off += k
            self._big_end, self._big_numel = (big_end if self._pm is not None else 0), big
            self._early_event = None
        flat, off = self.flat, n - 1
        pairs = [(v, p.grad.reshape(-1)) for p, v in zip(live, self._views)]
        todo = [(v, s) for v, s in pairs if s.data_ptr() != v.data_ptr()]          # (already in place: nothing to pack)
        early = self._early_event is not None
        if early and any(v.data_ptr() < flat.data_ptr() + 4 * self._big_end for v, _ in todo):
            raise RuntimeError("GradBucket: the large region was all-reduced early but a gradient of it is not in its slot")
        if todo:
            torch._foreach_copy_([v for v, _ in todo], [s for _, s in todo])
        # the flag rides in the same bucket; it is summed, so any rank's inf makes it non-zero everywhere
        # (device-side writes only: a Python scalar assignment is a host->device copy, which a CUDA-graph capture rejects)
        if found_inf is None:
            flat[off:off + 1].zero_()
        elif torch.is_tensor(found_inf):
            flat[off:off + 1].copy_(found_inf.reshape(1).to(torch.float32))
        else:
            flat[off:off + 1].fill_(float(found_inf))
        if ws > 1 and self._pm is not None:
            if early:
                main = torch.cuda.current_stream()
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    self._pm.all_reduce_(average, self._big_end, self._pm.n)   # small region + flag (+ the bucket's tail pad)
                main.wait_stream(self._side)
                self._early_event = None
            else:
                self._pm.all_reduce_(average)              # averaging is folded into the reduction kernel
        elif ws > 1:
            dist.all_reduce(flat, group=group)
            if average:
                flat[:off].mul_(1.0 / ws)
        flag = flat[off].clone()
        for p, v in zip(live, self._views):
            p.grad = v.view_as(p.grad)
        return flag

    def signature(self):
        """names-free description of which tensors are in the bucket (ranks must agree; checked by the tests)"""
        return [tuple(p.shape) for p in self._live()]


def cell_shard(n_cells, world_size, rank):
    """contiguous range of density-grid cells (per cascade) a rank evaluates during update_extra_state"""
    per = (n_cells + world_size - 1) // world_size
    lo = min(rank * per, n_cells)
    return lo, min(lo + per, n_cells)


def merge_density(fresh, group=None):
    """partial `fresh` grids (-1 where a rank evaluated nothing) -> identical merged grid on every rank"""
    ws, _ = world()
    if ws > 1:
        dist.all_reduce(fresh, op=dist.ReduceOp.MAX, group=group)
    return fresh


def shared_seed(seed_tensor=None, group=None):
    """broadcast rank 0's RNG seed so the stochastic cell selection of the density refresh matches on all ranks"""
    ws, rank = world()
    dev = "cpu"
    if ws > 1 and dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())      # NCCL moves device tensors only
    t = torch.zeros(1, dtype=torch.int64, device=dev) if seed_tensor is None else seed_tensor
    if rank == 0 and seed_tensor is None:
        t[0] = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    if ws > 1:
        dist.broadcast(t, src=0, group=group)
    return int(t.item())
