#!/usr/bin/env python
"""Skybox SH projection micro-benchmark (BASELINE configs[0] and the configs[4] sweep): device-
resident maps, CUDA events on the library's stream, distinct maps rotated so reads come from HBM.

  python tools/skybox_bench.py                    C1: 2048x1024 RGBA32F, L2, 8 maps rotated
  python tools/skybox_bench.py --sweep            C5: batched maps 512x256 .. 4096x2048, L2 and L3
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vlb = importlib.import_module("vulkan-light-bakery_b200")


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def time_launches(stream, fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--fmt", default="f32", choices=["f32", "u8"])
    ap.add_argument("--size", default="2048x1024")
    ap.add_argument("--maps", type=int, default=8)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = vlb.Context(0)
    ctx.set_stream(stream.cuda_stream)
    pk = peak()
    gen = torch.Generator(device=dev).manual_seed(1)

    def run(W, H, n_maps, order, fmt, batched):
        bpt = 16 if fmt == "f32" else 4
        if fmt == "f32":
            maps = torch.rand((n_maps, H, W, 4), device=dev, generator=gen, dtype=torch.float32)
        else:
            maps = torch.randint(0, 256, (n_maps, H, W, 4), device=dev, generator=gen, dtype=torch.uint8)
        outs = torch.zeros((n_maps, 48), device=dev)
        stride = H * W * bpt
        f = vlb.FMT_RGBA32F if fmt == "f32" else vlb.FMT_RGBA8
        if batched:
            us = time_launches(stream, lambda: ctx.skybox_project_sh_device(maps.data_ptr(), stride, n_maps, f, W, H, order, outs.data_ptr()), args.reps)
            bytes_ = stride * n_maps
        else:
            def fn():
                for i in range(n_maps):
                    ctx.skybox_project_sh_device(maps[i].data_ptr(), stride, 1, f, W, H, order, outs[i].data_ptr())
            us = time_launches(stream, fn, args.reps) / n_maps
            bytes_ = stride
        gbs = bytes_ / (us * 1e-6) / 1e9
        return {"W": W, "H": H, "maps": n_maps, "order": order, "fmt": fmt, "batched": batched, "us_per_launch": round(us, 3),
                "GBs": round(gbs, 1), "frac_measured_peak": round(gbs / pk, 4), "frac_8TBs": round(gbs / 8000.0, 4)}

    if not args.sweep:
        W, H = (int(x) for x in args.size.split("x"))
        print(json.dumps(run(W, H, args.maps, args.order, args.fmt, False)))
        print(json.dumps(run(W, H, args.maps, args.order, args.fmt, True)))
    else:
        for (W, H, n) in ((512, 256, 1024), (1024, 512, 1024), (2048, 1024, 256), (4096, 2048, 64)):
            for order in (2, 3):
                print(json.dumps(run(W, H, n, order, "f32", True)), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
