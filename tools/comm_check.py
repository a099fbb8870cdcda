#!/usr/bin/env python
"""One process per GPU through the library's own communicator (run under torchrun): vlb_comm_init_rank with the id
carried by torch.distributed, replicated uploads, vlb_bake_probes_sharded[_device] direct and multi-bounce; every rank
compares the gathered grid bit for bit with its own single-GPU bake of the whole grid. Prints one line per rank."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
vlb = importlib.import_module("vulkan-light-bakery_b200")
scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
uid = torch.zeros(vlb.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(vlb.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
ctx = vlb.Context(local)
ctx.comm_init_rank(uid.cpu().numpy().tobytes(), rank, world)
assert ctx.comm_info()[:2] == (rank, world)
scene = scenes.atrium(16384, seed=7); sky = scenes.hdr_sky(256, 128, seed=1)
ok = True
for probes, bounces in (((8, 4, 8), 0), ((8, 4, 7), 0), ((6, 4, 5), 2), ((4, 4, 1), 0)):
    s = scenes.atrium_settings(probes=probes, dirs=(32, 32), order=3, bounds=(0, 0, 0) + tuple(scenes.HALL))
    s.bounces, s.indirect_gain = bounces, 0.8
    ctx.comm_sharded_uploads(False)
    ctx.set_scene(scene); ctx.build_bvh(); ctx.set_skybox(sky)
    want = ctx.bake_probes(s)                          # whole grid on this GPU alone
    ctx.comm_sharded_uploads(True)                      # collective from here on
    ctx.set_skybox_async(sky); ctx.set_scene(scene); ctx.build_bvh()
    got = ctx.bake_probes_sharded(s, want_output=True)
    same = np.array_equal(got, want)
    if bounces == 0:
        full = torch.zeros((s.n_probes, 48), device=dev)
        ctx.bake_probes_sharded_device(s, 0, full.data_ptr()); ctx.synchronize()
        same = same and np.array_equal(full.cpu().numpy().reshape(-1, 16, 3), want)
    # own rows only: this rank's z-slices land in its rows of a host grid, the other rows are left alone
    grid = np.full((s.n_probes, 16, 3), -7.0, np.float32)
    ctx.bake_probes_sharded_rows(s, grid.ctypes.data)
    nxy = probes[0] * probes[1]
    g4 = grid.reshape(probes[2], nxy, 16, 3); w4 = want.reshape(probes[2], nxy, 16, 3)
    mine = np.zeros(probes[2], bool); mine[rank::world] = True
    same = same and np.array_equal(g4[mine], w4[mine]) and bool(np.all(g4[~mine] == -7.0))
    ok = ok and same
    print("rank %d/%d probes %s bounces %d: sharded == single-GPU bitwise: %s" % (rank, world, probes, bounces, same), flush=True)
ctx.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
