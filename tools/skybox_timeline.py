#!/usr/bin/env python
"""Per-CTA timeline of ONE k_project_tiles launch (C1: 2048x1024 RGBA32F -> L2 SH) from the %globaltimer stamps of
the instrumented build: VLB_BUILD_TAG=ptiming VLB_NVCC_EXTRA=-DVLB_PROJ_TIMING python -m ...build, then
VLB_LIB=.../libvlb_bake_ptiming.so VLB_PROJ_TIMING=1 python tools/skybox_timeline.py. Prints min/avg/max over the CTAs
(microseconds after the first CTA started) of: start, first tile landed, last tile consumed, strip flushed, partial
published, exit, tile 1 and tile 3 consumed."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vlb = importlib.import_module("vulkan-light-bakery_b200")
import torch
W, H = 2048, 1024
ctx = vlb.Context(0)
maps = torch.rand((6, H, W, 4), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
outs = torch.zeros((6, 48), device="cuda")
for rep in range(3):
    for i in range(4):
        flush.zero_(); torch.cuda.synchronize()
        ctx.skybox_project_sh_device(maps[i].data_ptr(), H * W * 16, 1, vlb.FMT_RGBA32F, W, H, 2, outs[i].data_ptr())
        ctx.synchronize()
        sys.stderr.write("single launch %d.%d (cold L2): " % (rep, i)); sys.stderr.flush()
        ctx.skybox_project_sh_device_ptrs([maps[5].data_ptr()], vlb.FMT_RGBA32F, W, H, 2, outs[5].data_ptr())   # dumps set 0
        ctx.synchronize()
