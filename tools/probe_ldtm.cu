// Probe: TMEM load cost per shape.  W warps each read their lane quarter (32 lanes x 64 columns = 8 KB of accumulators)
// `iters` times, either as 2 x tcgen05.ld.32x32b.x32 or as 2 x tcgen05.ld.16x256b.x8 (the fragment layout of the
// epilogue), with a tcgen05.wait::ld after each pair.  Reports cycles per 8 KB block and warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../convnet_quantization_b200/csrc/common.cuh"
namespace b200q { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } int launched(const char*) { return 0; } int num_sms() { return 148; }
int encode_tensor_map(CUtensorMap*, const void*, int, const uint64_t*, const uint64_t*, const uint32_t*, int) { return 0; } }
using namespace b200q;

template <int MODE>
__global__ void __launch_bounds__(512, 1) ldtm_probe(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) * 64 & 511);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t a[32], b[32];
    if (MODE == 0) {
      tmem_ld_32x32(base, a);
      tmem_ld_32x32(base + 32, b);
    } else {
      tmem_ld_16x256b_x8(base, a);
      tmem_ld_16x256b_x8(base + (16u << 16), b);
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= a[j] ^ b[j];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; uint32_t* sink;
  cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
  const int iters = 4096;
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 2; ++mode) {
      if (mode == 0) ldtm_probe<0><<<148, warps * 32>>>(iters, d, sink); else ldtm_probe<1><<<148, warps * 32>>>(iters, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      printf("%2d warps, %s: %.1f cycles per 8 KB block per warp, %.1f B/clk/SM [%s]\n", warps,
             mode == 0 ? "2 x 32x32b.x32" : "2 x 16x256b.x8", (double)c / iters, warps * 8192.0 * iters / c, cudaGetErrorString(e));
    }
  }
  return 0;
}
