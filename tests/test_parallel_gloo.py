"""The N > 1 path on CPU: one process per "GPU" (gloo, world_size 2 and 3), probes sharded by
z-slices (contiguous slabs and cyclic slices), one all-gather, result == the single-process grid.
The per-rank baker here is the CPU oracle (test infrastructure) or a synthetic function of the
global probe index; the sharding / gathering code under test is vulkan-light-bakery_b200/parallel.py."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, cyclic, probes, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    vlb = importlib.import_module("vulkan-light-bakery_b200")
    par = importlib.import_module("vulkan-light-bakery_b200.parallel")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = vlb.default_settings()
        s.probes[:] = probes
        s.dir_w, s.dir_h = 16, 8
        nxy = probes[0] * probes[1]
        if mode == "synthetic":
            def bake_slab(ss, out):
                ids = np.concatenate([np.arange(nxy) + k * nxy for k in ss.slab_slices]) if ss.slab_slices else np.zeros(0)
                out.copy_(torch.from_numpy((ids[:, None] * 100.0 + np.arange(48)[None, :]).astype(np.float32)))
        else:
            scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
            from oracle import oracle_api as oa
            oa.set_num_threads(2)
            osc = oa.Scene(scenes.small_room())
            s.flags = vlb.SHADOW_RAYS | vlb.SRGB_ENCODE
            s.light_pos[:] = (2.0, 3.5, 2.0)
            vlb.settings_from_bounds(s, osc.bounds(tight=True))

            def bake_slab(ss, out):
                got, _ = osc.bake_probes(ss)
                out.copy_(torch.from_numpy(got.reshape(-1, 48)))
        full = par.bake_sharded(bake_slab, s, rank, world, device="cpu", cyclic=cyclic)
        if rank == 0:
            ref = torch.empty((s.n_probes, 48))
            whole = s.copy()
            bake_slab(whole, ref)
            q.put(("ok", bool(torch.equal(full, ref)), tuple(full.shape)))
        # every rank must hold the same gathered buffer
        h = torch.tensor([float(full.double().sum())], dtype=torch.float64)
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        assert all(float(x) == float(hs[0]) for x in hs)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,cyclic,probes", [(2, False, (3, 2, 4)), (2, True, (3, 2, 4)), (3, False, (2, 2, 7)),
                                                (3, True, (2, 2, 7)), (2, True, (2, 1, 1))])
def test_sharded_bake_synthetic(world, cyclic, probes):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, "synthetic", cyclic, probes, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    tag, equal, shape = q.get(timeout=5)
    assert tag == "ok" and equal and shape == (probes[0] * probes[1] * probes[2], 48)


@pytest.mark.parametrize("cyclic", [False, True])
def test_sharded_bake_oracle_world2(cyclic):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, "oracle", cyclic, (2, 2, 3), q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    tag, equal, shape = q.get(timeout=5)
    assert tag == "ok" and equal and shape == (12, 48)


def _sky_worker(rank, world, port, n_maps, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    par = importlib.import_module("vulkan-light-bakery_b200.parallel")
    scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
    from oracle import oracle_api as oa
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def project(ids, out):
            for row, i in enumerate(ids):
                out[row].copy_(torch.from_numpy(oa.skybox_project(scenes.hdr_sky(32, 16, seed=100 + i), order=2).reshape(48)))
        full = par.project_maps_sharded(project, n_maps, rank, world, device="cpu")
        if rank == world - 1:
            ref = torch.empty((n_maps, 48))
            project(list(range(n_maps)), ref)
            q.put(("ok", bool(torch.equal(full, ref)), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_maps", [(2, 5), (3, 7), (3, 2), (2, 4)])
def test_sharded_skybox_batch(world, n_maps):
    """BASELINE configs[4] across GPUs: maps dealt round-robin, one all-gather of 192 bytes per map."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sky_worker, args=(r, world, port, n_maps, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    tag, equal, shape = q.get(timeout=5)
    assert tag == "ok" and equal and shape == (n_maps, 48)
