"""Multi-GPU bake: probes shard across the GPUs of one box by contiguous z-slabs of the grid (k is
the slowest index of i + j*Nx + k*Nx*Ny, so a slab is one contiguous byte range of the output),
scene + BVH + skybox replicated on every GPU, and the per-GPU SH slabs are gathered with ONE
all-gather (NCCL over NVLink on GPUs; gloo in the CPU tests). One process per GPU
(torch.distributed); there is no other data-path collective — probes are independent.

The reference is single-device (src/application.cpp:90-136 always picks physical device 0); this
is the build's only parallel axis.
"""
import numpy as np


def slab_range(nz, rank, world):
    """z-slices [k0, k1) of `rank`: contiguous, sizes differ by at most one, first ranks larger."""
    base, rem = divmod(int(nz), int(world))
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


def slab_sizes(nz, world):
    return [slab_range(nz, r, world)[1] - slab_range(nz, r, world)[0] for r in range(world)]


def shard_settings(settings, rank, world):
    s = settings.copy()
    s.slab_k0, s.slab_k1 = slab_range(settings.probes[2], rank, world)
    return s


def gather_slabs(local, settings, rank, world, group=None):
    """All-gathers the per-rank slabs ([n_local_probes, 48] torch tensors on the bake device) into
    the full [Nx*Ny*Nz, 48] buffer, identical on every rank. Slabs are padded to the largest slab
    so a single equal-size all_gather_into_tensor moves everything."""
    import torch
    import torch.distributed as dist
    nxy = settings.probes[0] * settings.probes[1]
    sizes = slab_sizes(settings.probes[2], world)
    pad = max(sizes) * nxy
    send = local
    if local.shape[0] != pad:
        send = torch.zeros((pad, 48), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    if world == 1:
        return local
    recv = torch.empty((world * pad, 48), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(sz == sizes[0] for sz in sizes):
        return recv
    parts = [recv[r * pad: r * pad + sizes[r] * nxy] for r in range(world)]
    return torch.cat(parts, 0)


def bake_sharded(bake_slab, settings, rank, world, device=None, group=None):
    """bake_slab(slab_settings, out_tensor) fills out_tensor ([n_local, 48] float32 on `device`)
    with this rank's slab; returns the gathered full grid on every rank."""
    import torch
    s = shard_settings(settings, rank, world)
    n_local = settings.probes[0] * settings.probes[1] * (s.slab_k1 - s.slab_k0)
    out = torch.empty((max(n_local, 1), 48), dtype=torch.float32, device=device)[:n_local]
    if n_local:
        bake_slab(s, out)
    return gather_slabs(out, settings, rank, world, group)
