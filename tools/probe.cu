// Micro-probes for design decisions (not part of the product): tcgen05 kind::i8 issue floor vs N, TMEM load rate,
// int<->float conversion rate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe tools/probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../convnet_quantization_b200/csrc/common.cuh"
namespace b200q { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } int launched(const char*) { return 0; } int num_sms() { return 148; }
int encode_tensor_map(CUtensorMap*, const void*, int, const uint64_t*, const uint64_t*, const uint32_t*, int) { return 0; } }
using namespace b200q;

// One CTA per SM; thread 0 issues `iters` MMAs (M=128, N, K=32 int8) on smem operands, commits, waits.
template <int N, int KC>
__global__ void __launch_bounds__(128, 1) mma_probe(int iters, long long* cycles, int shift_rows = 0, int tap_pattern = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a = smem;                 // 4 stages of A: 128 x KC (+ slack rows for shifted starts)
  uint8_t* b = smem + 4 * 128 * KC + 80 * KC;  // B: N x KC
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < (4 * 128 * KC + N * KC) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_i8(128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // tap_pattern: emulate the 9 row-shifted views of the halo kernel (shift = kh*33 + kw rows)
      const int sh = tap_pattern ? ((i % 9) / 3) * 33 + (i % 3) + shift_rows : shift_rows;
      const uint32_t a_addr = smem_u32(a + (tap_pattern ? 0 : (i & 3) * 128 * KC)) + sh * KC;
      const uint32_t b_addr = smem_u32(b);
#pragma unroll
      for (int k = 0; k < KC / 32; ++k)
        tc_mma_i8(tmem + (i & 1) * N, make_kmajor_desc<KC>(a_addr + k * 32, 8 * KC), make_kmajor_desc<KC>(b_addr + k * 32, 8 * KC), idesc, 1u);
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// Lean issue loop: descriptors precomputed, 8 MMAs per iteration back to back (no per-MMA integer work).
template <int N, int KC>
__global__ void __launch_bounds__(128, 1) mma_probe_lean(int iters, long long* cycles, int shift_rows = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a = smem;
  uint8_t* b = smem + 4 * 128 * KC;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < (4 * 128 * KC + N * KC) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x < 32) {
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc_i8(128, N);
    const uint64_t da0 = make_kmajor_desc<KC>(smem_u32(a) + shift_rows * KC, 8 * KC);
    const uint64_t db0 = make_kmajor_desc<KC>(smem_u32(b), 8 * KC);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
      if (leader) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          tc_mma_i8(tmem + (u & 1) * N, da0 + (uint64_t)(((u & 3) * 128 * KC + (u & 1) * 32) >> 4), db0 + (uint64_t)(((u & 1) * 32) >> 4), idesc, 1u);
      }
      __syncwarp();
    }
    if (leader) { tc_commit(&bar); }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int KC>
void run_lean(long long* d_cycles) {
  const int iters = 8192;
  const int smem = 4 * 128 * KC + N * KC + 2048;
  cudaFuncSetAttribute(mma_probe_lean<N, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int sh = 1; sh <= 8; ++sh) {
    mma_probe_lean<N, KC><<<148, 128, smem>>>(iters, d_cycles, sh);
    cudaDeviceSynchronize();
    long long cs = 0;
    cudaMemcpy(&cs, d_cycles, 8, cudaMemcpyDeviceToHost);
    printf("LEAN N=%3d KC=%3d A shifted %d rows: %.1f cyc/MMA\n", N, KC, sh, (double)cs / iters);
  }
  mma_probe_lean<N, KC><<<148, 128, smem>>>(iters, d_cycles, 0);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / iters;
  printf("LEAN mma i8 M=128 N=%3d KC=%3d: %.1f cyc/MMA (%.0f MAC/cyc/SM, smem operand %.0f B/cyc) [%s]\n", N, KC, per, 128.0 * N * 32 / per,
         (128.0 * 32 + N * 32.0) / per, cudaGetErrorString(e));
}

// TMEM load rate: W warps each load their lane quarter, 32 columns at a time, `iters` times.
__global__ void __launch_bounds__(512, 1) ldtm_probe(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int warp = threadIdx.x >> 5;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + ((i * 32 + (warp >> 2) * 64) & 511), v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc ^= v[j];
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// Conversion / ALU rates: MODE 0 = I2F+F2I, 1 = magic-number (IADD/FADD), 2 = FADD+FMUL, 3 = f32x2 add+mul
template <int MODE>
__global__ void __launch_bounds__(256) cvt_probe(int iters, long long* cycles, int* sink, int seed) {
  int v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = seed + threadIdx.x * 8 + j;
  const float m = 1.0009765625f, bd = 3.0f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if constexpr (MODE == 3) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        unsigned long long p, q, r;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "r"(v[j]), "r"(v[j + 1]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(bd), "f"(bd));
        asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(p), "l"(q));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(m), "f"(m));
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(r), "l"(q));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(v[j]), "=r"(v[j + 1]) : "l"(p));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if constexpr (MODE == 0) {
          float t = __int2float_rn(v[j]);
          v[j] = __float2int_rn(__fmul_rn(t, m)) + 1;
        } else if constexpr (MODE == 1) {
          float t = __fadd_rn(__int_as_float(v[j] + 0x4B400000), -12582912.0f);
          t = __fmul_rn(t, m);
          v[j] = __float_as_int(__fadd_rn(t, 12582912.0f)) - 0x4B400000 + 1;
        } else {
          float t = __fadd_rn(__int_as_float(v[j]), bd);
          v[j] = __float_as_int(__fmul_rn(t, m));
        }
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  int s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s ^= v[j];
  if (s == 0x7fffffff) sink[0] = s;
}

template <int N, int KC>
void run_mma(long long* d_cycles) {
  const int iters = 4096;
  const int smem = 4 * 128 * KC + 80 * KC + N * KC + 2048;
  cudaFuncSetAttribute(mma_probe<N, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (KC == 64) {
    for (int shift = 0; shift <= 9; ++shift) {
      mma_probe<N, KC><<<148, 128, smem>>>(iters, d_cycles, shift, 0);
      cudaDeviceSynchronize();
      long long c = 0;
      cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
      printf("mma i8 N=%3d KC=%3d A start shifted by %d rows: %.1f cyc/MMA\n", N, KC, shift, (double)c / (iters * (KC / 32)));
    }
    mma_probe<N, KC><<<148, 128, smem>>>(iters, d_cycles, 0, 1);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
    printf("mma i8 N=%3d KC=%3d 9-tap shifted pattern: %.1f cyc/MMA\n", N, KC, (double)c / (iters * (KC / 32)));
  }
  for (int grid : {1, 148}) {
    mma_probe<N, KC><<<grid, 128, smem>>>(iters, d_cycles);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * (KC / 32));
    printf("mma i8 M=128 N=%3d KC=%3d grid=%3d: %.1f cyc/MMA (%.0f MAC/cyc/SM) smem operand bytes/cyc=%.0f  [%s]\n", N, KC, grid, per,
           128.0 * N * 32 / per, (128.0 * 32 + N * 32.0) / per, cudaGetErrorString(e));
  }
}

int main() {
  long long* d_cycles; int* d_sink;
  cudaMalloc(&d_cycles, 64); cudaMalloc(&d_sink, 64);
  run_lean<16, 64>(d_cycles); run_lean<32, 64>(d_cycles); run_lean<64, 64>(d_cycles); run_lean<128, 64>(d_cycles); run_lean<64, 128>(d_cycles); run_lean<128, 128>(d_cycles); run_lean<256, 128>(d_cycles); run_lean<64, 32>(d_cycles);
  run_mma<64, 64>(d_cycles); run_mma<128, 64>(d_cycles); run_mma<64, 128>(d_cycles); run_mma<128, 128>(d_cycles); run_mma<256, 128>(d_cycles);
  run_mma<32, 64>(d_cycles); run_mma<16, 64>(d_cycles);
  for (int threads : {128, 256, 512}) {
    const int iters = 2048;
    ldtm_probe<<<148, threads>>>(iters, d_cycles, (uint32_t*)d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
    printf("ldtm 32x32b.x32: %d warps: %.1f cyc per (warp-load of 4 KB), %.1f B/cyc/SM [%s]\n", threads / 32, (double)c / iters,
           (double)(threads / 32) * 4096.0 * iters / c, cudaGetErrorString(e));
  }
  {
    const int iters = 4096; long long c;
    for (int blocks : {1, 2}) {
      cvt_probe<0><<<148 * blocks, 256>>>(iters, d_cycles, d_sink, 5); cudaDeviceSynchronize(); cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
      printf("I2F+FMUL+F2I+IADD x8/thread, %d warps/SM: %.2f cyc per warp-element-op-group\n", 8 * blocks, (double)c / (iters * 8.0));
      cvt_probe<1><<<148 * blocks, 256>>>(iters, d_cycles, d_sink, 5); cudaDeviceSynchronize(); cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
      printf("magic(IADD,FADD)+FMUL+FADD+IADD x8/thread, %d warps/SM: %.2f cyc\n", 8 * blocks, (double)c / (iters * 8.0));
      cvt_probe<2><<<148 * blocks, 256>>>(iters, d_cycles, d_sink, 5); cudaDeviceSynchronize(); cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
      printf("FADD+FMUL x8/thread, %d warps/SM: %.2f cyc\n", 8 * blocks, (double)c / (iters * 8.0));
      cvt_probe<3><<<148 * blocks, 256>>>(iters, d_cycles, d_sink, 5); cudaDeviceSynchronize(); cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
      printf("FADD2+FMUL2 (f32x2) x8/thread, %d warps/SM: %.2f cyc\n", 8 * blocks, (double)c / (iters * 8.0));
    }
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
