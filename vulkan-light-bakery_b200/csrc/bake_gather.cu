// bake_gather.cu -- the gather-pass instantiations of k_bake_stream (multi-bounce passes: shaders/main.rchit:124-167,
// shaders/sh.rmiss:20-36) as a translation unit of their own; see the head of bake.cu.
#define VLB_BAKE_GATHER_TU 1
#include "bake.cu"
