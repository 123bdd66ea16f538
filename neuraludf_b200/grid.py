"""Dense UDF grid queries for mesh extraction with a device-side hand-off (BASELINE config 5).

The reference fills an N^3 lattice by shipping every batch of points host -> device -> host and then evaluates the gradient
where `udf < 2 voxels`, again batch by batch through the host (`extract_mesh.py:18-105 get_udf_normals_grid_slow`).  Here
the lattice coordinates are generated on the device, the value sweep runs on the fused value-chain kernel, the
near-surface cells are compacted on the device and only they go through the fused value + reverse-sweep kernel; the result
is handed over as the dense distance grid plus a SPARSE list of near-surface cells (flat index + unit normal) -- 67 MB +
~7 MB at 256^3 instead of the reference's 67 MB + 201 MB dense normal grid and its ~130 blocking copies.
`get_udf_normals_grid_slow` keeps the reference's return convention for `udf_mc_lewiner` (the Cython MeshUDF marching cubes,
out of scope)."""
import torch
import torch.nn.functional as F


def lattice_points(head, count, N, device):
    """points `head .. head+count-1` of the N^3 lattice on [-1,1]^3 in the reference's order (x slowest, z fastest;
    extract_mesh.py:38-51), generated on the device"""
    voxel = 2.0 / (N - 1)
    idx = torch.arange(head, head + count, device=device, dtype=torch.int64)
    k = idx % N
    j = torch.div(idx, N, rounding_mode="floor") % N
    i = torch.div(torch.div(idx, N, rounding_mode="floor"), N, rounding_mode="floor") % N
    return torch.stack([i.float() * voxel - 1.0, j.float() * voxel - 1.0, k.float() * voxel - 1.0], dim=-1)


@torch.no_grad()
def udf_grid(udf_network, N, max_batch=1 << 21, lo=0, hi=None):
    """udf at the lattice points [lo, hi) (default: all N^3) as a flat device tensor -- one fused-chain launch per batch,
    no host round trips.  Slab partitioning for multi-GPU sweeps: give each rank its own [lo, hi)."""
    device = next(udf_network.parameters()).device
    hi = N ** 3 if hi is None else hi
    out = torch.empty(hi - lo, device=device)
    for head in range(lo, hi, max_batch):
        n = min(max_batch, hi - head)
        out[head - lo:head - lo + n] = udf_network.udf_values(lattice_points(head, n, N, device))
    return out


@torch.no_grad()
def near_surface_cells(udf_network, N, df_flat=None, max_batch=1 << 20, dist_voxels=2.0, lo=0):
    """(flat lattice indices [M] int64, unit vectors pointing towards the surface [M,3]) of the cells with
    udf < dist_voxels * voxel (extract_mesh.py:77-98): compaction and gradient evaluation stay on the device."""
    device = next(udf_network.parameters()).device
    if df_flat is None:
        df_flat = udf_grid(udf_network, N)
    voxel = 2.0 / (N - 1)
    idx = torch.nonzero(df_flat < dist_voxels * voxel).reshape(-1) + lo
    normals = torch.empty(idx.numel(), 3, device=device)
    for head in range(0, idx.numel(), max_batch):
        sel = idx[head:head + max_batch]
        k = sel % N
        j = torch.div(sel, N, rounding_mode="floor") % N
        i = torch.div(torch.div(sel, N, rounding_mode="floor"), N, rounding_mode="floor") % N
        pts = torch.stack([i.float() * voxel - 1.0, j.float() * voxel - 1.0, k.float() * voxel - 1.0], dim=-1)
        g = udf_network.gradient(pts)[:, 0]
        g = g / (torch.linalg.norm(g, ord=2, dim=-1, keepdim=True) + 1e-5)          # exp_runner_blending.py:767-771 (func_grad)
        normals[head:head + sel.numel()] = -F.normalize(g, dim=1)                    # extract_mesh.py:93
    return idx, normals


@torch.no_grad()
def get_udf_normals_grid_slow(udf_network, N=56, max_batch=1 << 20):
    """Same returns as the reference's function of this name (df_values [N,N,N], vecs [N,N,N,3], samples [N^3,7], on the
    host) -- assembled on the device, copied once."""
    device = next(udf_network.parameters()).device
    df = udf_grid(udf_network, N, max_batch)
    idx, normals = near_surface_cells(udf_network, N, df, max(max_batch // 4, 1))
    samples = torch.zeros(N ** 3, 7, device=device)
    samples[:, 0:3] = lattice_points(0, N ** 3, N, device)
    samples[:, 3] = df
    samples[idx, 4:] = normals
    samples = samples.cpu()
    return samples[:, 3].reshape(N, N, N), samples[:, 4:].reshape(N, N, N, 3), samples
