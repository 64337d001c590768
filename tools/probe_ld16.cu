// Probe: register <-> (TMEM lane, column) mapping of tcgen05.ld.16x256b.x8 (used by the pooled epilogue: a thread
// must hold vertically adjacent accumulator rows).  Every TMEM cell is tagged (lane << 8 | column) through 32x32b
// stores, read back through 16x256b.x8 at lane offsets +0 and +16 of each warp's quarter, and compared on the host
// with the expected fragment layout  row = 16*half + t/4 + 8*((r>>1)&1),  col = 8*(r>>2) + 2*(t&3) + (r&1).
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../convnet_quantization_b200/csrc/common.cuh"
namespace b200q { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } int launched(const char*) { return 0; } int num_sms() { return 148; }
int encode_tensor_map(CUtensorMap*, const void*, int, const uint64_t*, const uint64_t*, const uint32_t*, int) { return 0; } }
using namespace b200q;

__global__ void __launch_bounds__(128, 1) ld16_probe(uint32_t* out) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t row_addr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = ((uint32_t)(warp * 32 + lane) << 8) | (uint32_t)(c0 + i);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(row_addr + c0),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  tmem_st_wait();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    tmem_ld_16x256b_x8(row_addr + ((uint32_t)(half * 16) << 16), r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[((warp * 2 + half) * 32 + lane) * 32 + i] = r[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 4 * 2 * 32 * 32 * 4);
  ld16_probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<uint32_t> h(4 * 2 * 32 * 32);
  cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w)
    for (int half = 0; half < 2; ++half)
      for (int t = 0; t < 32; ++t)
        for (int r = 0; r < 32; ++r) {
          const uint32_t got = h[((w * 2 + half) * 32 + t) * 32 + r];
          const uint32_t row = 32 * w + 16 * half + t / 4 + 8 * ((r >> 1) & 1), col = 8 * (r >> 2) + 2 * (t & 3) + (r & 1);
          if (got != ((row << 8) | col)) {
            if (bad < 40) printf("w%d half%d t%2d r%2d: got lane %u col %u, expected lane %u col %u\n", w, half, t, r, got >> 8, got & 255, row, col);
            ++bad;
          }
        }
  printf("ld16 probe: %d mismatches of %zu\n", bad, h.size());
  for (int r = 0; r < 8; ++r) printf("  warp0 half0 t5 r%d -> lane %u col %u\n", r, h[5 * 32 + r] >> 8, h[5 * 32 + r] & 255);
  return bad != 0;
}
