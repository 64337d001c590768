// Probe: tcgen05.mma kind::i8 cycles per MMA (M=128, K=32) when the A operand's 8-row groups are NOT contiguous
// (descriptor SBO = image-row pitch of the halo kernels) and when the start address is shifted by whole rows, i.e.
// the access patterns of conv_halo.cu.  One CTA per SM, one elected thread issues back-to-back MMAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../convnet_quantization_b200/csrc/common.cuh"
namespace b200q { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } int launched(const char*) { return 0; } int num_sms() { return 148; }
int encode_tensor_map(CUtensorMap*, const void*, int, const uint64_t*, const uint64_t*, const uint32_t*, int) { return 0; } }
using namespace b200q;

template <int N, int KC>
__global__ void __launch_bounds__(128, 1) sbo_probe(int iters, long long* cycles, int shift_rows, int sbo_bytes, int taps, int pitch_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a = smem;                 // up to 80 KB of A
  uint8_t* b = smem + 96 * 1024;     // B: 9 x N x KC
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < (96 * 1024 + 9 * N * KC) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x < 32) {
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc_i8(128, N);
    const uint64_t da0 = make_kmajor_desc<KC>(smem_u32(a) + shift_rows * KC, sbo_bytes);
    const uint64_t db0 = make_kmajor_desc<KC>(smem_u32(b), 8 * KC);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 9 * (KC / 32)) {
      if (leader) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int sh = taps ? (tap / 3) * pitch_rows + (tap % 3) : 0;
#pragma unroll
          for (int k = 0; k < KC / 32; ++k)
            tc_mma_i8(tmem + ((i / 9) & 1) * N, da0 + (uint64_t)((sh * KC + k * 32) >> 4), db0 + (uint64_t)((tap * N * KC + k * 32) >> 4), idesc, 1u);
        }
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

// Chains of `chain` MMAs per accumulator, rotating over `slots` TMEM accumulators: cost of switching the accumulator,
// of the per-chain tcgen05.commit, and of starting a chain with accumulate=1 (pre-biased accumulator) vs 0.
template <int N, int KC>
__global__ void __launch_bounds__(128, 1) chain_probe(int chains, long long* cycles, int chain, int slots, int commit_each, int first_acc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a = smem;
  uint8_t* b = smem + 96 * 1024;
  __shared__ uint64_t bar, bars[8];
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < (96 * 1024 + N * KC) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(bars + i, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x < 32) {
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc_i8(128, N);
    const uint64_t da0 = make_kmajor_desc<KC>(smem_u32(a), 8 * KC);
    const uint64_t db0 = make_kmajor_desc<KC>(smem_u32(b), 8 * KC);
    long long t0 = clock64();
    for (int c = 0; c < chains; ++c) {
      if (leader) {
        const uint32_t d = tmem + (c % slots) * N;
        for (int m = 0; m < chain; ++m)
          tc_mma_i8(d, da0 + (uint64_t)(((m & 15) * 1024) >> 4), db0, idesc, (m == 0) ? (uint32_t)first_acc : 1u);
        if (commit_each) tc_commit(bars + (c & 7));
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
void run_chain(long long* d, int chain, int slots, int commit_each, int first_acc) {
  const int chains = 1024;
  const int smem = 96 * 1024 + N * KC + 2048;
  cudaFuncSetAttribute(chain_probe<N, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  chain_probe<N, KC><<<148, 128, smem>>>(chains, d, chain, slots, commit_each, first_acc);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("CHAIN N=%3d KC=%3d chain=%3d slots=%d commit_each=%d first_acc=%d: %.1f cyc/MMA, %.0f cyc/chain [%s]\n", N, KC, chain, slots,
         commit_each, first_acc, (double)c / (chains * (double)chain), (double)c / chains, cudaGetErrorString(e));
}

template <int N, int KC>
void run(long long* d, const char* what, int shift, int sbo, int taps, int pitch) {
  const int iters = 9 * (KC / 32) * 512;
  const int smem = 96 * 1024 + 9 * N * KC + 2048;
  cudaFuncSetAttribute(sbo_probe<N, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  sbo_probe<N, KC><<<148, 128, smem>>>(iters, d, shift, sbo, taps, pitch);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d KC=%3d shift=%2d sbo=%5d taps=%d pitch=%2d: %.1f cyc/MMA  %s [%s]\n", N, KC, shift, sbo, taps, pitch, (double)c / iters, what,
         cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  for (int chain : {18, 36, 72}) {
    run_chain<64, 64>(d, chain, 1, 0, 1);
    run_chain<64, 64>(d, chain, 4, 0, 1);
    run_chain<64, 64>(d, chain, 4, 1, 1);
    run_chain<64, 64>(d, chain, 4, 1, 0);
    run_chain<128, 64>(d, chain, 1, 0, 1);
    run_chain<128, 64>(d, chain, 4, 0, 1);
    run_chain<128, 64>(d, chain, 4, 1, 1);
    run_chain<128, 64>(d, chain, 4, 1, 0);
    run_chain<128, 64>(d, chain, 2, 1, 0);
  }
  run_chain<256, 64>(d, 18, 2, 1, 1);
  run_chain<256, 64>(d, 18, 2, 1, 0);
  run_chain<256, 64>(d, 72, 2, 1, 1);
  run_chain<256, 64>(d, 72, 2, 1, 0);
  run<64, 64>(d, "canonical", 0, 512, 0, 0);
  run<64, 64>(d, "start +1 row", 1, 512, 0, 0);
  run<64, 64>(d, "start +2 rows", 2, 512, 0, 0);
  run<64, 64>(d, "sbo = 33 rows", 0, 33 * 64, 0, 0);
  run<64, 64>(d, "sbo = 34 rows", 0, 34 * 64, 0, 0);
  run<64, 64>(d, "sbo = 36 rows", 0, 36 * 64, 0, 0);
  run<64, 64>(d, "sbo = 40 rows", 0, 40 * 64, 0, 0);
  run<64, 64>(d, "conv2 pattern (pitch 33)", 0, 33 * 64, 1, 33);
  run<64, 64>(d, "conv2 pattern (pitch 34)", 0, 34 * 64, 1, 34);
  run<64, 64>(d, "conv2 pattern (pitch 36)", 0, 36 * 64, 1, 36);
  run<64, 64>(d, "conv2 pattern (pitch 40)", 0, 40 * 64, 1, 40);
  run<128, 64>(d, "canonical", 0, 512, 0, 0);
  run<128, 64>(d, "conv3 pattern (pitch 17)", 0, 17 * 64, 1, 17);
  run<128, 64>(d, "conv3 pattern (pitch 18)", 0, 18 * 64, 1, 18);
  run<128, 64>(d, "conv3 pattern (pitch 20)", 0, 20 * 64, 1, 20);
  run<128, 64>(d, "conv3 pattern (pitch 24)", 0, 24 * 64, 1, 24);
  run<128, 128>(d, "canonical", 0, 1024, 0, 0);
  run<128, 128>(d, "conv4 pattern (pitch 17)", 0, 17 * 128, 1, 17);
  run<128, 128>(d, "conv4 pattern (pitch 18)", 0, 18 * 128, 1, 18);
  run<128, 128>(d, "conv4 pattern (pitch 24)", 0, 24 * 128, 1, 24);
  run<64, 128>(d, "canonical", 0, 1024, 0, 0);
  run<256, 128>(d, "canonical", 0, 1024, 0, 0);
  run<256, 128>(d, "conv5 pattern (pitch 9, M=2 images)", 0, 9 * 128, 1, 9);
  return 0;
}
