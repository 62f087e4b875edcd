"""`_backend` for raymarching: the reference's pybind11 module surface (raymarching/src/bindings.cpp:5-23), same
function names and argument order, bound to the C ABI of libpnerf_b200.so. Tensors are caller-allocated CUDA
tensors; nothing here allocates or synchronises."""
import ctypes

import torch

from .. import _lib as L
from .._lib import ptr, stream, call, require_cuda


L.lib.pnerf_occupied_bounds_floats.restype = ctypes.c_uint32
OCC_FLOATS = int(L.lib.pnerf_occupied_bounds_floats())


class _Backend:
    @staticmethod
    def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
        require_cuda(rays_o, rays_d, aabb, nears, fars)
        call("pnerf_near_far_from_aabb", ptr(rays_o), ptr(rays_d), ptr(aabb), N, min_near, ptr(nears), ptr(fars), stream())

    @staticmethod
    def sph_from_ray(rays_o, rays_d, radius, N, coords):
        require_cuda(rays_o, rays_d, coords)
        call("pnerf_sph_from_ray", ptr(rays_o), ptr(rays_d), radius, N, ptr(coords), stream())

    @staticmethod
    def morton3D(coords, N, indices):
        require_cuda(coords, indices)
        call("pnerf_morton3D", ptr(coords), N, ptr(indices), stream())

    @staticmethod
    def morton3D_invert(indices, N, coords):
        require_cuda(coords, indices)
        call("pnerf_morton3D_invert", ptr(indices), N, ptr(coords), stream())

    @staticmethod
    def packbits(grid, N, density_thresh, bitfield):
        require_cuda(grid, bitfield)
        call("pnerf_packbits", ptr(grid), N, density_thresh, ptr(bitfield), stream())

    @staticmethod
    def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas,
                         rays, counter, noises):
        require_cuda(rays_o, rays_d, grid, nears, fars, xyzs, dirs, deltas, rays, counter, noises)
        call("pnerf_march_rays_train", ptr(rays_o), ptr(rays_d), ptr(grid), bound, dt_gamma, max_steps, N, C, H, M,
             ptr(nears), ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(counter), ptr(noises), stream())

    @staticmethod
    def march_rays_train_ws(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas,
                            rays, counter, noises, t_list, occ_aabb, valid_rows=None):
        """one-walk variant (not in the reference): t_list [N, max_steps] scratch, occ_aabb [6] from occupied_bounds or None,
        valid_rows: optional int32 [1] receiving the number of leading sample rows that were written"""
        require_cuda(rays_o, rays_d, grid, nears, fars, xyzs, dirs, deltas, rays, counter, noises, t_list, occ_aabb, valid_rows)
        if t_list.numel() < N * max_steps or t_list.dtype != torch.float32:
            raise RuntimeError("t_list must hold N * max_steps floats")
        call("pnerf_march_rays_train_ws", ptr(rays_o), ptr(rays_d), ptr(grid), bound, dt_gamma, max_steps, N, C, H, M,
             ptr(nears), ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(counter), ptr(noises), ptr(t_list),
             ptr(occ_aabb), ptr(valid_rows), stream())

    @staticmethod
    def occupied_bounds(bitfield, C, H, bound, occ_aabb):
        require_cuda(bitfield, occ_aabb)
        if bitfield.numel() < C * H * H * H // 8 or occ_aabb.numel() < OCC_FLOATS or occ_aabb.dtype != torch.float32:
            raise RuntimeError(f"occupied_bounds: bitfield [C*H^3/8] uint8, occ_aabb [{OCC_FLOATS}] float32 (6 bounds + scratch)")
        call("pnerf_occupied_bounds", ptr(bitfield), C, H, float(bound), ptr(occ_aabb), stream())

    @staticmethod
    def composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image):
        require_cuda(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        call("pnerf_composite_rays_train_forward", ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(rays), M, N, T_thresh,
             ptr(weights_sum), ptr(depth), ptr(image), stream())

    @staticmethod
    def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N,
                                      T_thresh, grad_sigmas, grad_rgbs):
        require_cuda(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, grad_sigmas, grad_rgbs)
        call("pnerf_composite_rays_train_backward", ptr(grad_weights_sum), ptr(grad_image), ptr(sigmas), ptr(rgbs),
             ptr(deltas), ptr(rays), ptr(weights_sum), ptr(image), M, N, T_thresh, ptr(grad_sigmas), ptr(grad_rgbs),
             stream())

    @staticmethod
    def composite_rays_flex_train_forward(sigmas, input, deltas, rays, M, N, n_channel, T_thresh, output):
        require_cuda(sigmas, input, deltas, rays, output)
        call("pnerf_composite_rays_flex_train_forward", ptr(sigmas), ptr(input), ptr(deltas), ptr(rays), M, N, n_channel,
             T_thresh, ptr(output), stream())

    @staticmethod
    def composite_rays_flex_train_backward(grad_output, sigmas, input, deltas, rays, output, M, N, n_channel, T_thresh,
                                           grad_input):
        require_cuda(grad_output, sigmas, input, deltas, rays, output, grad_input)
        call("pnerf_composite_rays_flex_train_backward", ptr(grad_output), ptr(sigmas), ptr(input), ptr(deltas),
             ptr(rays), ptr(output), M, N, n_channel, T_thresh, ptr(grad_input), stream())

    @staticmethod
    def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears,
                   fars, xyzs, dirs, deltas, noises):
        require_cuda(rays_alive, rays_t, rays_o, rays_d, grid, nears, fars, xyzs, dirs, deltas, noises)
        call("pnerf_march_rays", n_alive, n_step, ptr(rays_alive), ptr(rays_t), ptr(rays_o), ptr(rays_d), bound, dt_gamma,
             max_steps, C, H, ptr(grid), ptr(nears), ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(noises), stream())

    @staticmethod
    def composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image):
        require_cuda(rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image)
        call("pnerf_composite_rays", n_alive, n_step, T_thresh, ptr(rays_alive), ptr(rays_t), ptr(sigmas), ptr(rgbs),
             ptr(deltas), ptr(weights_sum), ptr(depth), ptr(image), stream())

    @staticmethod
    def composite_rays_flex(n_alive, n_step, n_channel, T_thresh, rays_alive, rays_t, sigmas, input, deltas, weights_sum,
                            output):
        require_cuda(rays_alive, rays_t, sigmas, input, deltas, weights_sum, output)
        call("pnerf_composite_rays_flex", n_alive, n_step, n_channel, T_thresh, ptr(rays_alive), ptr(rays_t), ptr(sigmas),
             ptr(input), ptr(deltas), ptr(weights_sum), ptr(output), stream())

    @staticmethod
    def spread_ray_to_sample(input, rays, M, N, n_channel, output):
        require_cuda(input, rays, output)
        call("pnerf_spread_ray_to_sample", ptr(input), ptr(rays), M, N, n_channel, ptr(output), stream())


_backend = _Backend()
