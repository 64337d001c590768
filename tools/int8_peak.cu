// Measured int8 tensor-pipe rate of this GPU (roofline denominator): every SM issues back-to-back
// tcgen05.mma.kind::i8 (M=128, K=32) on shared-memory operands; the rate is total ops / CUDA-event time, and the SM
// clock the run held is read from clock64 / globaltimer inside the kernel.  Prints one JSON object.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/int8_peak tools/int8_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../convnet_quantization_b200/csrc/common.cuh"
namespace b200q {
void set_error(const char*, ...) {}
int check_cuda(cudaError_t, const char*) { return 0; }
int launched(const char*) { return 0; }
int num_sms() { return 148; }
bool pdl_enabled() { return false; }
void pdl_set(bool) {}
void note_graph_replay(int) {}
int encode_tensor_map(CUtensorMap*, const void*, int, const uint64_t*, const uint64_t*, const uint32_t*, int) { return 0; }
}  // namespace b200q
using namespace b200q;

template <int N>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, unsigned long long* cyc_ns) {
  constexpr int KC = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a = smem;                 // 4 x [128][KC]
  uint8_t* b = smem + 4 * 128 * KC;  // [N][KC]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < (4 * 128 * KC + N * KC) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_i8(128, N);
    uint64_t da[4], db[4];
    for (int k = 0; k < 4; ++k) {
      da[k] = make_kmajor_desc<KC>(smem_u32(a) + k * 32, 8 * KC);
      db[k] = make_kmajor_desc<KC>(smem_u32(b) + k * 32, 8 * KC);
    }
    const long long c0 = clock64();
    const uint64_t t0 = globaltimer_ns();
    for (int i = 0; i < iters; ++i) {
      const uint64_t st = (uint64_t)(((i & 3) * 128 * KC) >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_i8(tmem + (i & 1) * N, da[k] + st, db[k], idesc, 1u);
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    const long long c1 = clock64();
    const uint64_t t1 = globaltimer_ns();
    if (blockIdx.x == 0) {
      cyc_ns[0] = (unsigned long long)(c1 - c0);
      cyc_ns[1] = (unsigned long long)(t1 - t0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N>
static void run(int sms, int iters, bool last) {
  auto k = peak_kernel<N>;
  const int smem = 4 * 128 * 128 + N * 128 + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  unsigned long long* d;
  cudaMalloc(&d, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<<<sms, 128, smem>>>(1000, d);  // warm-up
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<<<sms, 128, smem>>>(iters, d);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double mmas = 4.0 * iters;
  const double mac_per_clk = 128.0 * N * 32 * mmas / (double)h[0];
  const double tops = 2.0 * 128 * N * 32 * mmas * sms / (ms * 1e-3) / 1e12;
  printf("    {\"n\": %d, \"mma_per_sm\": %.0f, \"cycles_per_mma\": %.2f, \"mac_per_clk_per_sm\": %.1f, \"sm_mhz_during_run\": %.0f, "
         "\"ms\": %.3f, \"tops\": %.1f, \"error\": \"%s\"}%s\n",
         N, mmas, (double)h[0] / mmas, mac_per_clk, 1e3 * (double)h[0] / (double)h[1], ms, tops,
         cudaGetErrorString(cudaGetLastError()), last ? "" : ",");
  cudaFree(d);
}

int main(int argc, char** argv) {
  int dev = 0, sms = 148, mhz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, dev);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, dev);
  const int burst = 200000;                              // ~10-25 ms per run: a burst, like a 40 ms bench window
  const int sustained = argc > 1 ? atoi(argv[1]) : 16000000;  // ~1-2 s: under the power cap
  printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %d,\n", p.name, sms, mhz / 1000);
  printf("  \"what\": \"tcgen05.mma.cta_group::1.kind::i8 M=128 K=32, operands in shared memory (SW128), one issuing thread per SM, "
         "all SMs; ops = 2*M*N*K per MMA; tops from CUDA events around the launch\",\n");
  printf("  \"burst\": [\n");
  run<64>(sms, burst, false);
  run<128>(sms, burst, false);
  run<256>(sms, burst, true);
  printf("  ],\n  \"sustained\": [\n");
  run<128>(sms, sustained, false);
  run<256>(sms, sustained / 2, true);
  printf("  ]\n}\n");
  return 0;
}
