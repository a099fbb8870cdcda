"""Committed fixtures of the hot path (tests/golden/oracle_golden.npz, made by tests/golden/make_oracle_golden.py):
CPU: the oracle still reproduces them (no silent drift of the specification; only libm's last ulp may differ
between machines, hence 1e-6). GPU (-m gpu): the CUDA path, through the C ABI, against the same committed numbers
with BASELINE.json's tolerances -- 1e-4 per skybox SH vector, 1e-3 per probe SH vector."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "oracle_golden.npz")


def _gen():
    spec = importlib.util.spec_from_file_location("make_oracle_golden", os.path.join(HERE, "golden", "make_oracle_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_oracle_reproduces_the_committed_fixtures(gold):
    now = _gen().compute()
    assert sorted(now) == sorted(gold.files)
    for k in gold.files:
        assert rel_l2(now[k], gold[k]) <= 1e-6, k


@pytest.mark.gpu
def test_cuda_path_against_committed_fixtures(ctx, vlb, scenes, gold):
    g = _gen()
    for seed, (w, h) in ((1, (64, 32)), (2, (100, 37))):
        sky = scenes.hdr_sky(w, h, seed=seed)
        for order in (2, 3):
            assert rel_l2(ctx.skybox_project_sh(sky, order=order), gold["skybox_f32_s%d_o%d" % (seed, order)]) <= 1e-4
            assert rel_l2(ctx.envmap_project_sh(sky, order=order), gold["envmap_f32_s%d_o%d" % (seed, order)]) <= 1e-4
        u8 = (np.clip(sky, 0, 1) * 255).astype(np.uint8)
        assert rel_l2(ctx.skybox_project_sh(u8, order=3), gold["skybox_u8_s%d_o3" % seed]) <= 1e-4
    sky = scenes.hdr_sky(64, 32, seed=1)
    for name, sc in (("room", scenes.small_room()), ("room_textured", scenes.small_room_textured())):
        ctx.set_scene(sc)
        ctx.build_bvh()
        ctx.set_skybox(sky)
        for order in (2, 3):
            assert rel_l2(ctx.bake_probes(g.bake_settings(order)), gold["bake_%s_o%d" % (name, order)]) <= 1e-3
    ctx.set_scene(scenes.small_room())
    ctx.build_bvh()
    s = g.bake_settings(3, bounces=1)
    assert rel_l2(ctx.bake_probes(s), gold["gather_room_pass1"]) <= 1e-3
