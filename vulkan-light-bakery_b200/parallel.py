"""Multi-GPU bake: probes shard across the GPUs of one box by z-slices of the grid (k is the slowest
index of i + j*Nx + k*Nx*Ny, so a slice is one contiguous byte range of the output), scene + BVH +
skybox replicated on every GPU, and the per-GPU SH buffers are gathered with ONE all-gather (NCCL
over NVLink on GPUs; gloo in the CPU tests). One process per GPU (torch.distributed); there is no
other data-path collective -- probes are independent.

Two partitions of the slices:
  contiguous  rank r bakes one slab [k0, k1) (sizes differ by at most one, first ranks larger)
  cyclic      rank r bakes k = r, r + world, r + 2*world, ...  Every rank then samples the whole depth
              of the scene, which evens out the cost of probes near geometry; the gathered buffer
              is un-interleaved with one strided device copy. This is what bench.py uses.

The reference is single-device (src/application.cpp:90-136 always picks physical device 0); this
is the build's only parallel axis (SURVEY \u00a78e).
"""
import numpy as np


def slab_range(nz, rank, world):
    """z-slices [k0, k1) of `rank`: contiguous, sizes differ by at most one, first ranks larger."""
    base, rem = divmod(int(nz), int(world))
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


def slab_sizes(nz, world):
    return [slab_range(nz, r, world)[1] - slab_range(nz, r, world)[0] for r in range(world)]


def cyclic_slices(nz, rank, world):
    return list(range(rank, int(nz), int(world)))


def shard_settings(settings, rank, world, cyclic=False):
    s = settings.copy()
    if cyclic:
        s.slab_k0, s.slab_k1, s.slab_stride = rank, settings.probes[2], world
    else:
        s.slab_k0, s.slab_k1 = slab_range(settings.probes[2], rank, world)
        s.slab_stride = 1
    return s


def order_after_bake(ctx):
    """Stream contract of this module. The collectives below are issued on torch's CURRENT stream, a vlb Context by
    default works on its own non-blocking stream: unless the two are the same stream, the all-gather could read the
    slab before k_bake_stream has written it (and the next gather pass a stale `prev`). Pass the Context as `ctx`
    to the functions below: if it was put on torch's current stream (ctx.set_stream(torch.cuda.current_stream().cuda_stream),
    what bench.py does) nothing more is needed, otherwise the bake is waited for here."""
    if ctx is None:
        return
    import torch
    if not torch.cuda.is_available():
        return
    cur = torch.cuda.current_stream().cuda_stream
    if getattr(ctx, "stream_handle", None) != cur:
        ctx.synchronize()


def gather_slabs(local, settings, rank, world, group=None, cyclic=False, ctx=None):
    """All-gathers the per-rank buffers ([n_local_probes, 48] torch tensors on the bake device) into
    the full [Nx*Ny*Nz, 48] buffer in x-fastest order, identical on every rank. Buffers are padded to
    the largest share so a single equal-size all_gather_into_tensor moves everything.
    `ctx`: the vlb Context that baked `local` (see order_after_bake)."""
    import torch
    import torch.distributed as dist
    order_after_bake(ctx)
    if world == 1:
        return local
    nz = settings.probes[2]
    nxy = settings.probes[0] * settings.probes[1]
    sizes = [len(cyclic_slices(nz, r, world)) for r in range(world)] if cyclic else slab_sizes(nz, world)
    pad = max(sizes) * nxy
    send = local
    if local.shape[0] != pad:
        send = torch.zeros((pad, 48), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    recv = torch.empty((world * pad, 48), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if cyclic:
        # recv[r, i] is slice k = r + i*world: transpose (rank, i) -> (i, rank) and drop the padding slices
        full = recv.view(world, max(sizes), nxy, 48).transpose(0, 1).reshape(max(sizes) * world * nxy, 48)
        return full[: nz * nxy] if full.shape[0] != nz * nxy else full
    if all(sz == sizes[0] for sz in sizes):
        return recv
    parts = [recv[r * pad: r * pad + sizes[r] * nxy] for r in range(world)]
    return torch.cat(parts, 0)


def bake_sharded(bake_slab, settings, rank, world, device=None, group=None, cyclic=False, ctx=None):
    """bake_slab(slab_settings, out_tensor) fills out_tensor ([n_local, 48] float32 on `device`)
    with this rank's share; returns the gathered full grid on every rank. `ctx`: see order_after_bake."""
    import torch
    s = shard_settings(settings, rank, world, cyclic)
    n_local = s.n_slab_probes
    out = torch.empty((max(n_local, 1), 48), dtype=torch.float32, device=device)[:n_local]
    if n_local:
        bake_slab(s, out)
    return gather_slabs(out, settings, rank, world, group, cyclic, ctx)


def bake_multibounce_sharded(bake_pass, settings, rank, world, device=None, group=None, cyclic=False, ctx=None):
    """Multi-bounce bake of a sharded grid (include/vlb_bake.h: vlb_bake_gather_device). Every pass needs
    the PREVIOUS pass over the whole grid, so this is the one place on the path with a real exchange
    step: 1 + settings.bounces passes, each followed by one all-gather of the slabs.
    bake_pass(slab_settings, prev_full_or_None, out_tensor) fills out_tensor ([n_local, 48]) with this
    rank's share of the pass; prev_full is the gathered [n_probes, 48] tensor of the pass before.
    Returns the gathered last pass on every rank."""
    import torch
    s = shard_settings(settings, rank, world, cyclic)
    n_local = s.n_slab_probes
    prev = None
    for _ in range(1 + max(0, int(settings.bounces))):
        out = torch.empty((max(n_local, 1), 48), dtype=torch.float32, device=device)[:n_local]
        if n_local:
            bake_pass(s, prev, out)
        full = gather_slabs(out, settings, rank, world, group, cyclic, ctx)
        prev = full.contiguous()
    return prev


# ---- batched skybox projection (BASELINE configs[4]) ------------------------------------------------------
def map_share(n_maps, rank, world):
    """Maps of `rank` under the round-robin deal: rank, rank + world, ..."""
    return list(range(int(rank), int(n_maps), int(world)))


def project_maps_sharded(project, n_maps, rank, world, device=None, group=None, ctx=None):
    """Batched skybox SH projection over `world` GPUs: maps are independent, so they are dealt round-robin,
    every rank projects its share with project(map_ids, out) (out: [len(map_ids), 48] float32 on `device`,
    e.g. Context.skybox_project_sh_device over its resident maps) and ONE all-gather of the padded shares gives
    every rank all n_maps x 48 coefficients in map order. No other collective: 192 bytes per map."""
    import torch
    import torch.distributed as dist
    mine = map_share(n_maps, rank, world)
    out = torch.zeros((max(len(mine), 1), 48), dtype=torch.float32, device=device)[: len(mine)]
    if mine:
        project(mine, out)
    order_after_bake(ctx)
    if world == 1:
        return out
    pad = (int(n_maps) + world - 1) // world
    send = out
    if out.shape[0] != pad:
        send = torch.zeros((pad, 48), dtype=torch.float32, device=device)
        send[: out.shape[0]] = out
    recv = torch.empty((world * pad, 48), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    # recv[r, i] is map r + i * world
    full = recv.view(world, pad, 48).transpose(0, 1).reshape(pad * world, 48)
    return full[: int(n_maps)].contiguous()
