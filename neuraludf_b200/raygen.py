"""Ray generation on the device with the call shapes of the reference's data loader.

The reference builds every batch of rays with a dozen small torch ops on the host-driven stream
(`dataset/dataset.py:228-294 gen_random_rays_patches_at`, `:151-164 gen_rays_at`, `:329-335 near_far_from_sphere`).
These functions take the reference's `Dataset` object (only its tensors are read: `images`, `masks`, `intrinsics_all_inv`,
`pose_all`, `H`, `W`) and do the same work in ONE libnudf kernel per call: pixel -> camera ray -> world ray, colour / mask
gather, normalised uv and the unit-sphere near / far bounds.  The random pixel draw stays in torch (same RNG stream as the
reference).  The patch crop of the fine-tuning stage (`crop_patch=True`, one `grid_sample`) is kept as the reference has it.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib as L


def _cam(dataset, img_idx):
    idx = int(img_idx)
    ki = dataset.intrinsics_all_inv[idx, :3, :3].contiguous().float()
    pose = dataset.pose_all[idx].contiguous().float()
    return ki, pose


def near_far_from_sphere(rays_o, rays_d):
    """dataset/dataset.py:329-335 (torch ops; the fused kernels below return near / far directly)."""
    a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return mid - 1.0, mid + 1.0


def gen_random_rays_patches_at(dataset, img_idx, batch_size, importance_sample=False, h_patch_size=3, crop_patch=False,
                               with_near_far=False):
    """Same dict as the reference's method (`rays`, `rays_ndc_uv`, `rays_norm_XYZ_cam`, `rays_patch_color`, `rays_patch_mask`)
    plus, on request, `near` / `far` [N,1]."""
    if importance_sample:
        raise NotImplementedError("importance_sample=True is never used by exp_runner_blending.py")
    lib = L.lib()
    dev = dataset.images.device
    W, H = int(dataset.W), int(dataset.H)
    px = torch.randint(low=0, high=W, size=[batch_size], device=dev)      # same two draws, same order as the reference
    py = torch.randint(low=0, high=H, size=[batch_size], device=dev)
    ki, pose = _cam(dataset, img_idx)
    idx = int(img_idx)
    image = dataset.images[idx].contiguous()
    mask = dataset.masks[idx].contiguous()
    rays = torch.empty(batch_size, 10, device=dev)
    uv = torch.empty(batch_size, 2, device=dev)
    near = torch.empty(batch_size, 1, device=dev) if with_near_far else None
    far = torch.empty(batch_size, 1, device=dev) if with_near_far else None
    with torch.cuda.device(dev):
        L.check(lib.nudf_gen_rays(L.ptr(ki), L.ptr(pose), L.ptr(px), L.ptr(py), batch_size, L.ptr(image), L.ptr(mask), H, W,
                                  L.ptr(rays), L.ptr(uv), L.ptr(near), L.ptr(far), L.stream_ptr()), "nudf_gen_rays")
    patch_color, patch_mask = None, None
    if crop_patch:                                           # dataset.py:252-266
        offs = torch.arange(-h_patch_size, h_patch_size + 1, device=dev)
        offsets = torch.stack(torch.meshgrid(offs, offs, indexing="ij")[::-1], dim=-1).view(1, -1, 2)
        grid = torch.stack([px, py], dim=-1).view(-1, 1, 2) + offsets.float()
        patch_mask = ((px > h_patch_size) * (px < (W - h_patch_size)) * (py > h_patch_size) * (py < H - h_patch_size)).view(-1, 1)
        guv = torch.stack([2 * grid[:, :, 0] / (W - 1) - 1, 2 * grid[:, :, 1] / (H - 1) - 1], dim=-1)
        patch_color = F.grid_sample(image[None].permute(0, 3, 1, 2), guv[None], mode="bilinear", padding_mode="zeros",
                                    align_corners=False)[0].permute(1, 2, 0).contiguous()
    p = torch.stack([px, py, torch.ones_like(py)], dim=-1).float()
    sample = {"rays": rays, "rays_ndc_uv": uv, "rays_norm_XYZ_cam": torch.matmul(ki[None], p[:, :, None]).squeeze(-1),
              "rays_patch_color": patch_color, "rays_patch_mask": patch_mask}
    if with_near_far:
        sample["near"], sample["far"] = near, far
    return sample


def gen_rays_at(dataset, img_idx, resolution_level=1, with_near_far=False):
    """(rays_o, rays_d) of shape [H // l, W // l, 3] like the reference's method; with_near_far adds ([.., 1], [.., 1])."""
    lib = L.lib()
    dev = dataset.images.device
    W, H = int(dataset.W), int(dataset.H)
    Wl, Hl = W // resolution_level, H // resolution_level
    ki, pose = _cam(dataset, img_idx)
    rays_o = torch.empty(Hl, Wl, 3, device=dev)
    rays_d = torch.empty(Hl, Wl, 3, device=dev)
    near = torch.empty(Hl, Wl, 1, device=dev) if with_near_far else None
    far = torch.empty(Hl, Wl, 1, device=dev) if with_near_far else None
    with torch.cuda.device(dev):
        L.check(lib.nudf_gen_rays_grid(L.ptr(ki), L.ptr(pose), W, H, Wl, Hl, L.ptr(rays_o), L.ptr(rays_d), L.ptr(near), L.ptr(far),
                                       L.stream_ptr()), "nudf_gen_rays_grid")
    return (rays_o, rays_d, near, far) if with_near_far else (rays_o, rays_d)
