"""The multi-GPU side of the C ABI (include/vlb_bake.h "multi-GPU", csrc/comm.cu): NCCL communicator per ctx, sharded
bake + all-gather, replicated uploads. On a one-GPU box the tests run with a one-rank communicator and the
no-communicator route; the two-rank cases need >= 2 GPUs (gpurun --gpus 2) and are skipped otherwise. The
one-process-per-GPU route is checked by tools/comm_check.py under torchrun (tools/r2_multi.sh)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _settings(vlb, bounces=0):
    s = vlb.default_settings()
    s.probes[:] = (3, 2, 5); s.dir_w, s.dir_h = 32, 16; s.sh_order = 3; s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE
    vlb.settings_from_bounds(s, (0.3, 0.3, 0.3, 3.7, 3.7, 3.7))
    s.bounces, s.indirect_gain = bounces, 0.7
    return s


@pytest.fixture(scope="module")
def room(scenes):
    return scenes.small_room()


def test_sharded_bake_without_and_with_a_one_rank_communicator(vlb, scenes, room):
    import torch
    sc = room
    sky = scenes.hdr_sky(64, 32, seed=2)
    with vlb.Context(0) as c:
        c.set_scene(sc); c.build_bvh(); c.set_skybox(sky)
        for bounces in (0, 2):
            s = _settings(vlb, bounces)
            want = c.bake_probes(s)
            assert np.array_equal(c.bake_probes_sharded(s), want)          # no communicator: the whole grid on this ctx
        assert c.comm_info()[:2] == (0, 1)
        with pytest.raises(vlb.VlbError):
            c.comm_sharded_uploads(True)                                    # needs a communicator
        c.comm_init_rank(vlb.comm_unique_id(), 0, 1)
        assert c.comm_info()[:2] == (0, 1) and c.comm_info()[2] >= 22000
        with pytest.raises(vlb.VlbError):
            c.comm_init_rank(vlb.comm_unique_id(), 0, 1)                    # already has one
        c.comm_sharded_uploads(True)
        c.set_scene(sc); c.build_bvh(); c.set_skybox(sky)                   # one rank: plain copies
        s = _settings(vlb)
        want = c.bake_probes(s)
        assert np.array_equal(c.bake_probes_sharded(s), want)
        full = torch.zeros((s.n_probes, 48), device="cuda")
        c.bake_probes_sharded_device(s, 0, full.data_ptr()); c.synchronize()
        assert np.array_equal(full.cpu().numpy().reshape(-1, 16, 3), want)
        t = s.copy(); t.slab_k0, t.slab_k1 = 0, 2
        with pytest.raises(vlb.VlbError):
            c.bake_probes_sharded(t)                                        # takes the whole grid
        c.comm_destroy()
        assert c.comm_info()[:2] == (0, 1)


def test_two_ranks_in_one_process(vlb, scenes, room):
    """vlb_comm_init_all + vlb_bake_probes_multi over NCCL (distinct devices): bit-identical to one GPU, direct and
    multi-bounce; replicated uploads give the same scene on both GPUs."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sc = room
    sky = scenes.hdr_sky(64, 32, seed=2)
    sky8 = (np.clip(sky, 0, 1) * 255).astype(np.uint8)
    ctxs = [vlb.Context(0), vlb.Context(1)]
    try:
        for c in ctxs:
            c.set_scene(sc); c.build_bvh(); c.set_skybox(sky)
        want = {b: ctxs[0].bake_probes(_settings(vlb, b)) for b in (0, 2)}
        for b in (0, 2):                                                    # first call creates the communicator
            assert np.array_equal(vlb.bake_probes_multi(ctxs, _settings(vlb, b)), want[b]), b
        assert [c.comm_info()[:2] for c in ctxs] == [(0, 2), (1, 2)]
        # replicated uploads: collective, so one thread per rank
        for sk in (sky, sky8):
            want_sk = None
            with vlb.Context(0) as ref:
                ref.set_scene(sc); ref.build_bvh(); ref.set_skybox(sk)
                want_sk = ref.bake_probes(_settings(vlb))
            errs = []

            def upload(c):
                try:
                    c.comm_sharded_uploads(True)
                    c.set_scene(sc); c.build_bvh(); c.set_skybox(sk)
                    c.comm_sharded_uploads(False)
                except Exception as e:      # noqa
                    errs.append(e)
            th = [threading.Thread(target=upload, args=(c,)) for c in ctxs]
            [t.start() for t in th]; [t.join() for t in th]
            assert not errs, errs
            for c in ctxs:
                assert np.array_equal(c.bake_probes(_settings(vlb)), want_sk)
            assert np.array_equal(vlb.bake_probes_multi(ctxs, _settings(vlb)), want_sk)
    finally:
        for c in ctxs:
            c.close()
