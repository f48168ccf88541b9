// Single-pass digit engine (ozaki5.cuh), host entry points used by tb_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
struct tb_gp;
namespace tb {
int oz5_init();                              // kernel attributes (once per process)
int oz5_ensure(tb_gp* gp);                   // (re)build row stats + digit tiles after a cache refresh; sets gp->oz5_mode
int oz5_tile_width(const tb_gp* gp);         // candidates per tile in the chosen mode
size_t oz5_tile_bytes(const tb_gp* gp);      // K* digit bytes per candidate tile
int oz5_launch_kstar(tb_gp* gp, cudaStream_t st, const double* Xc_dev, int64_t mc, int tiles, int8_t* BS, double* mean);
int oz5_launch_gemm(tb_gp* gp, cudaStream_t st, const int8_t* BS, int tiles, int G, int64_t McPad, double* partial);
}  // namespace tb
