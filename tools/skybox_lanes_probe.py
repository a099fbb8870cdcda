import ctypes, importlib, os, sys, time, torch
sys.path.insert(0, '/root/repo')
vlb = importlib.import_module("vulkan-light-bakery_b200")
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx = vlb.Context(0)
ctx.set_stream(stream.cuda_stream)
fmt, bpt = vlb.FMT_RGBA32F, 16
pad = int(os.environ.get("PAD", "0"))
n = 1024*2048*4
big = torch.rand(8 * (n + pad // 4), device=dev)
maps = [big[i * (n + pad // 4): i * (n + pad // 4) + n] for i in range(8)]
outs2 = torch.zeros((8, 48), device=dev)
stride = n * 4
ptrs = (ctypes.c_void_p * 8)(*[m.data_ptr() for m in maps])
def chain():
    ctx.skybox_project_sh_device_ptrs(ptrs, fmt, 2048, 1024, 2, outs2.data_ptr())
chain(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for r in range(50): chain()
b.record(stream)
torch.cuda.synchronize()
us = a.elapsed_time(b)*1e3/(400)
print(os.environ.get("TAG", ""), "pad", pad, "chained gpu us/launch %.2f GB/s %.0f" % (us, stride/us/1e3))
