// Probe: can a UMMA K-major swizzled smem descriptor start at an arbitrary ROW (pixel) of a tile that TMA wrote?
// A tile [ROWS][SW bytes] is laid out exactly as TMA SWIZZLE_{128,64}B would (address-bit XOR, tile base 1024-aligned);
// one MMA (M=128, N=64, K=32) with B = "identity" extracts D[i][n] = A[i + shift][kslice*32 + n]; host checks it.
// Variants: descriptor base_offset field = 0 or (start_addr >> 7) & 7.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../convnet_quantization_b200/csrc/common.cuh"
namespace b200q { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } int launched(const char*) { return 0; } int num_sms() { return 148; }
int encode_tensor_map(CUtensorMap*, const void*, int, const uint64_t*, const uint64_t*, const uint32_t*, int) { return 0; } }
using namespace b200q;

template <int SW>
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t byte) {  // offset of (row, byte) inside a 1024-aligned tile
  const uint32_t lin = row * SW + byte;
  if (SW == 128) return lin ^ (((lin >> 7) & 7) << 4);
  if (SW == 64) return lin ^ (((lin >> 7) & 3) << 4);
  return lin ^ (((lin >> 7) & 1) << 4);
}

template <int SW>
__global__ void __launch_bounds__(128, 1) shift_probe(int shift, int kslice, int use_base_offset, int32_t* out) {
  constexpr int ROWS = 192;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a = smem;                    // [ROWS][SW]
  uint8_t* b = smem + ROWS * 128;       // [64][SW]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < ROWS * SW; i += blockDim.x) {
    const int r = i / SW, k = i % SW;
    a[swz<SW>(r, k)] = (uint8_t)((r * 7 + k * 3 + 1) & 0xFF);
  }
  for (int i = threadIdx.x; i < 64 * SW; i += blockDim.x) {
    const int n = i / SW, k = i % SW;
    b[swz<SW>(n, k)] = (uint8_t)((n < 32 && (k % 32) == n) ? 1 : 0);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_i8(128, 64);
    const uint32_t a_addr = smem_u32(a) + shift * SW + kslice * 32;
    const uint32_t b_addr = smem_u32(b) + kslice * 32;
    uint64_t da = make_kmajor_desc<SW>(a_addr, 8 * SW);
    if (use_base_offset) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t db = make_kmajor_desc<SW>(b_addr, 8 * SW);
    tc_mma_i8(tmem, da, db, idesc, 0u);
    tc_commit(&bar);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  tc_fence_after();
  {
    const int warp = threadIdx.x >> 5;
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[threadIdx.x * 32 + j] = (int32_t)v[j];
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

template <int SW>
void run(int32_t* d_out) {
  const int smem = 192 * 128 + 64 * 128 + 2048;
  cudaFuncSetAttribute(shift_probe<SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<int32_t> h(128 * 32);
  for (int ks = 0; ks < SW / 32; ++ks)
    for (int ubo = 0; ubo < 2; ++ubo) {
      printf("SW%-3d kslice %d base_offset=%s: mismatches per shift:", SW, ks, ubo ? "(addr>>7)&7" : "0");
      for (int shift = 0; shift <= 17; ++shift) {
        shift_probe<SW><<<1, 128, smem>>>(shift, ks, ubo, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" [%s]", cudaGetErrorString(e)); break; }
        cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 128; ++i)
          for (int n = 0; n < 32; ++n) bad += h[i * 32 + n] != (((i + shift) * 7 + (ks * 32 + n) * 3 + 1) & 0xFF);
        printf(" %d:%d", shift, bad);
      }
      printf("\n");
    }
}

int main() {
  int32_t* d_out; cudaMalloc(&d_out, 128 * 32 * 4);
  run<128>(d_out); run<64>(d_out); run<32>(d_out);
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
