"""Generates tests/golden/oracle_golden.npz: outputs of the CPU oracle (oracle/vlb_oracle.cpp) on small seeded
inputs of every hot-path entry point -- skybox / env-map projection, direct bake, textured bake, gather pass. The
reference ships no vectors of its own and cannot run here, so these fixtures do not pin the oracle to
the reference (tests/golden/sh_common_golden.npz does that for what is compilable); they freeze the oracle, so that
neither it nor the CUDA path can drift unnoticed, and they let the GPU box compare against committed numbers.

    python tests/golden/make_oracle_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_api as oa  # noqa: E402

vlb = importlib.import_module("vulkan-light-bakery_b200")
scenes = importlib.import_module("vulkan-light-bakery_b200.scenes")


def bake_settings(order=3, bounces=0):
    s = vlb.default_settings()
    s.probes[:] = (3, 2, 3)
    s.dir_w, s.dir_h = 32, 16
    s.sh_order = order
    s.light_pos[:] = (2.0, 3.5, 2.0)
    s.flags = vlb.SHADOW_RAYS | vlb.SKYBOX_ON_MISS | vlb.SRGB_ENCODE
    s.bounces, s.indirect_gain = bounces, 0.7
    vlb.settings_from_bounds(s, (0.3, 0.3, 0.3, 3.7, 3.7, 3.7))
    return s


def compute():
    out = {}
    for seed, (w, h) in ((1, (64, 32)), (2, (100, 37))):
        sky = scenes.hdr_sky(w, h, seed=seed)
        for order in (2, 3):
            out["skybox_f32_s%d_o%d" % (seed, order)] = oa.skybox_project(sky, order=order)
            out["envmap_f32_s%d_o%d" % (seed, order)] = oa.envmap_project(sky, order=order)
        u8 = (np.clip(sky, 0, 1) * 255).astype(np.uint8)
        out["skybox_u8_s%d_o3" % seed] = oa.skybox_project(u8, order=3)
    sky = scenes.hdr_sky(64, 32, seed=1)
    for name, sc in (("room", scenes.small_room()), ("room_textured", scenes.small_room_textured())):
        o = oa.Scene(sc)
        o.set_skybox(sky)
        for order in (2, 3):
            out["bake_%s_o%d" % (name, order)], _ = o.bake_probes(bake_settings(order))
    o = oa.Scene(scenes.small_room())
    o.set_skybox(sky)
    s = bake_settings(3)
    direct, _ = o.bake_probes(s)
    out["gather_room_pass1"], _ = o.bake_gather(s, direct)
    return out


def main():
    out = compute()
    path = os.path.join(ROOT, "tests", "golden", "oracle_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
