// Probe: how fast can a warp run the epilogue's requantisation arithmetic by itself (no TMEM, no barriers, stores to a
// private scratch line)?  W warps per SM each process `iters` "half blocks" (32 accumulators per thread, 16 channels,
// two pixels) with epilogue16.cuh's epi_half (+ 32 integer adds that refresh the inputs).  Measured on B200: 249 cycles
// with one warp per scheduler, 475 / 713 / 950 with two / three / four: the scheduler retires one half block per ~240
// cycles however many warps share it, i.e. the arithmetic (64 packed FP + 48 integer-pipe instructions per half block)
// is pipe-bound, not latency-bound.  That is the floor under conv1 (32 blocks per image) and conv3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../convnet_quantization_b200/csrc/common.cuh"
#include "../convnet_quantization_b200/csrc/epilogue16.cuh"
namespace b200q { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } int launched(const char*) { return 0; } int num_sms() { return 148; }
int encode_tensor_map(CUtensorMap*, const void*, int, const uint64_t*, const uint64_t*, const uint32_t*, int) { return 0; } }
using namespace b200q;

struct Consts { int32_t cm[64]; float k1[64], bdiv[64], mult[64]; };

__global__ void __launch_bounds__(512, 1) requant_probe(const __grid_constant__ Consts consts, int iters, long long* cycles,
                                                        uint8_t* scratch, uint32_t seed) {
  __shared__ uint32_t magic;
  if (threadIdx.x == 0) magic = MAGIC_BITS;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int ch0 = 16 * (lane & 3);
  EpiRegs<16> K;
  epi_init(consts, ch0, &magic, K);
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = MAGIC_BITS + ((seed * (threadIdx.x + 1) * (i + 3)) & 0xffff);
  uint8_t* out = scratch + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 64;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    epi_half<false, 16>(v, K, consts, ch0, true, 0, 0, out, 16, true, true);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += 3 + it;  // new inputs every iteration (dependent on nothing computed above)
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
  long long* d; uint8_t* scratch;
  cudaMalloc(&d, 8); cudaMalloc(&scratch, 148 * 512 * 64);
  Consts c;
  for (int i = 0; i < 64; ++i) { c.cm[i] = (int32_t)MAGIC_BITS; c.k1[i] = -MAGIC_F - (float)i; c.bdiv[i] = 0.25f * i; c.mult[i] = 0.001f * (i + 1); }
  const int iters = 4096;
  for (int warps : {1, 4, 8, 12, 16}) {
    requant_probe<<<148, warps * 32>>>(c, iters, d, scratch, 12345u);
    cudaError_t e = cudaDeviceSynchronize();
    long long cy = 0; cudaMemcpy(&cy, d, 8, cudaMemcpyDeviceToHost);
    printf("%2d warps/SM: %.1f cycles per half block (32 values/thread) per warp [%s]\n", warps, (double)cy / iters, cudaGetErrorString(e));
  }
  return 0;
}
