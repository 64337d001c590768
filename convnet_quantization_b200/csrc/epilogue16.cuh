// Epilogue shared by the block-tile tensor-core kernels (conv1_tc.cu, conv_halo.cu): TMEM -> exact fbgemm
// requantisation (+ fused 2x2 max-pool) -> uint8 NHWC.
//
// The first version read the accumulators with tcgen05.ld.32x32b (thread = one pixel, 16 channels at a time, constants
// through the uniform datapath) and pooled AFTER requantising with warp shuffles on packed bytes.  ncu showed those
// layers issue-bound in the epilogue (conv2: 230 instructions per 32x16 accumulator block, conv1: 78 % issue-active), so
// this version is organised around the instruction count instead:
//
// * Accumulators are read with tcgen05.ld.16x256b (mma fragment layout, tools/probe_ld16.cu): a thread holds rows
//   t/4 and t/4 + 8 of a 16-lane half and the column pairs 8i + 2(t&3) + {0,1}.  In the kernels' 8-column x 16-row
//   pixel blocks (accumulator row = 8*image_row + column) rows 8 apart are VERTICALLY ADJACENT pixels, so the
//   vertical half of the 2x2 max-pool is one VIMNMX per value inside a thread, on the raw accumulators, before any
//   requantisation arithmetic; the horizontal half is one SHFL.BFLY(4) per pooled value, arranged so that the two
//   threads of a pair end up with different pooled rows (no duplicated work).  Pooled layers requantise 1/4 of the
//   values.
// * A thread always works on the same 16 output channels, so k1/bdiv/mult live in 48 registers for the lifetime of the
//   CTA: no constant loads in the loop.
// * The weight rows (B operand) are stored in shared memory in a permuted order so that the 16 channels of a thread
//   are CONSECUTIVE output channels (epi16_channel_of_column): one 16-byte store per pixel and thread, 64 contiguous
//   bytes per pixel and 4-thread group.
// * A warp takes a whole 32-row x 64-column block (one TMEM lane quarter, one 64-channel part) of a tile; warps are
//   grouped in sets that take alternate tiles (accumulator slots), so per-tile overhead (address arithmetic, barrier
//   wait, slot hand-back) is paid once per 2048 accumulators instead of once per 512.
// * Pre-biased accumulators (common.cuh requant4_prebiased) are re-armed with tcgen05.st from 8 long-lived registers.
#pragma once
#include "common.cuh"

namespace b200q {

// TMEM column (= B operand row) -> output channel, within each 64-channel part:
//   column = 16u + 8b + 2q + e   holds   channel = 16q + 4u + 2b + e        (u,q in 0..3; b,e in 0..1)
__host__ __device__ constexpr int epi16_channel_of_column(int col) {
  return (col & ~63) | (16 * ((col >> 1) & 3) + 4 * ((col >> 4) & 3) + 2 * ((col >> 3) & 1) + (col & 1));
}

struct Epi16Regs {
  float k1[16], bd[16], mu[16];  // per-thread constants of output channels ch0 .. ch0+15
  uint32_t fill[8];              // MAGIC_BITS x 8, opaque to the compiler so that it keeps them in registers
};

// ch0 = 64*part + 16*(lane & 3)
template <class Consts>
__device__ __forceinline__ void epi16_init(const Consts& c, int ch0, Epi16Regs& K) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    K.k1[k] = c.k1[ch0 + k];
    K.bd[k] = c.bdiv[ch0 + k];
    K.mu[k] = c.mult[ch0 + k];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) asm volatile("mov.u32 %0, 0x4B400000;" : "=r"(K.fill[i]));
}

// 16 lanes x 32 columns (fragment layout, see tmem_ld_16x256b_x8): register r = 8u + 4b + 2s + e of thread t holds
// lane base + t/4 + 8s, column col + 16u + 8b + 2(t&3) + e.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 16 columns: register r = 4b + 2s + e holds lane base + t/4 + 8s, column col + 8b + 2(t&3) + e.
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// Re-arm 16 lanes x 16 columns with the pre-bias.
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&f)[8]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(f[0]),
               "r"(f[1]), "r"(f[2]), "r"(f[3]), "r"(f[4]), "r"(f[5]), "r"(f[6]), "r"(f[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&f)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(f[0]),
               "r"(f[1]), "r"(f[2]), "r"(f[3]), "r"(f[4]), "r"(f[5]), "r"(f[6]), "r"(f[7])
               : "memory");
}

// Four consecutive channels K0..K0+3 (compile-time offset into the thread's 16) of one pixel -> packed word.
template <bool CHECK, int K0>
__device__ __forceinline__ uint32_t epi16_requant4(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3,
                                                   const Epi16Regs& K, int zp_sub, int lo, uint32_t& bad) {
  return requant4_prebiased<CHECK>(v0, v1, v2, v3, make_float4(K.k1[K0], K.k1[K0 + 1], K.k1[K0 + 2], K.k1[K0 + 3]),
                                   make_float4(K.bd[K0], K.bd[K0 + 1], K.bd[K0 + 2], K.bd[K0 + 3]),
                                   make_float4(K.mu[K0], K.mu[K0 + 1], K.mu[K0 + 2], K.mu[K0 + 3]), zp_sub, lo, bad);
}
// Exact conversion-pipe form of the same (rare: |acc| >= 2^22, or constants not flagged B200Q_RQ_BOUNDED).
template <class Consts>
__device__ __noinline__ uint32_t epi16_requant4_exact(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3,
                                                      const Consts& c, int ch, int zp_out, int lo) {
  // v = raw + MAGIC_BITS (wrapping)  ->  raw - corr
  const uint32_t v[4] = {v0, v1, v2, v3};
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int acc = (int)(v[i] + (uint32_t)c.cm[ch + i] - 2u * MAGIC_BITS);
    out |= requant_u8(acc, c.bdiv[ch + i], c.mult[ch + i], zp_out, lo) << (8 * i);
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------- no pooling
// One warp, one 32-lane quarter x 64 columns of an accumulator slot.  t_addr = tmem base + (quarter*32 << 16) + first
// column.  Thread (j = lane/4, q = lane&3) produces channels ch0..ch0+15 of the pixels (row0 + i, col0 + j), i = 0..3,
// of the tile block row this quarter covers; out = address of pixel (row0, col0 + j) channel ch0, row_stride = bytes
// between image rows.  `release` is called (by all lanes, converged) once the slot has been read and re-armed.
template <bool CHECK, class Consts, class Release>
__device__ __forceinline__ void epi16_block(uint32_t t_addr, const Epi16Regs& K, const Consts& consts, int ch0, bool fast,
                                            int zp_out, int lo, uint8_t* out, int64_t row_stride, bool valid,
                                            Release release) {
  const int zp_sub = zp_out - (int)MAGIC_BITS;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t a = t_addr + ((uint32_t)(16 * half) << 16);
    uint32_t v[32];
    tmem_ld_16x256b_x8(a, v);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 64; c += 16) tmem_st_16x256b_x2(a + c, K.fill);
    if (half == 1) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      release();
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {  // the thread's two pixels of this half: block rows 2*half + s
      uint32_t packed[4];
      uint32_t bad = 0;
      if (fast) {
        packed[0] = epi16_requant4<CHECK, 0>(v[2 * s], v[2 * s + 1], v[4 + 2 * s], v[5 + 2 * s], K, zp_sub, lo, bad);
        packed[1] = epi16_requant4<CHECK, 4>(v[8 + 2 * s], v[9 + 2 * s], v[12 + 2 * s], v[13 + 2 * s], K, zp_sub, lo, bad);
        packed[2] = epi16_requant4<CHECK, 8>(v[16 + 2 * s], v[17 + 2 * s], v[20 + 2 * s], v[21 + 2 * s], K, zp_sub, lo, bad);
        packed[3] = epi16_requant4<CHECK, 12>(v[24 + 2 * s], v[25 + 2 * s], v[28 + 2 * s], v[29 + 2 * s], K, zp_sub, lo, bad);
      }
      if (!fast || (CHECK && __any_sync(0xffffffffu, requant_magic_out_of_range(bad)))) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          packed[g] = epi16_requant4_exact(v[8 * g + 2 * s], v[8 * g + 2 * s + 1], v[8 * g + 4 + 2 * s],
                                           v[8 * g + 5 + 2 * s], consts, ch0 + 4 * g, zp_out, lo);
      }
      if (valid)
        *reinterpret_cast<uint4*>(out + (2 * half + s) * row_stride) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------- 2x2 max-pool
// Same block; the quarter's 4 image rows x 8 columns become 2 x 4 pooled pixels.  Thread (j, q) ends up with the pooled
// pixel (pooled row j&1, pooled column j>>1) of the block, channels ch0..ch0+15; out = its address.
// max() on the biased bit patterns is monotone in the raw accumulator (|acc| < 2^27), and the requantisation is
// monotone non-decreasing in the accumulator, so pooling first equals aten::quantized_max_pool2d on the stored tensor.
template <bool CHECK, class Consts, class Release>
__device__ __forceinline__ void epi16_block_pool(uint32_t t_addr, const Epi16Regs& K, const Consts& consts, int ch0,
                                                 bool fast, int zp_out, int lo, uint8_t* out, bool valid, int lane,
                                                 Release release) {
  const int zp_sub = zp_out - (int)MAGIC_BITS;
  const bool odd = (lane & 4) != 0;
  uint32_t packed[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {  // 16 columns at a time = channels 4u .. 4u+3 of the thread (keeps the live set small)
    const uint32_t a0 = t_addr + 16 * u, a1 = a0 + (16u << 16);
    uint32_t top[8], bot[8];
    tmem_ld_16x256b_x2(a0, top);  // block rows 0,1
    tmem_ld_16x256b_x2(a1, bot);  // block rows 2,3
    tmem_ld_wait();
    tmem_st_16x256b_x2(a0, K.fill);
    tmem_st_16x256b_x2(a1, K.fill);
    if (u == 3) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      release();
    }
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // k = 2b + e  <-  registers 4b + 2s + e
      const int i0 = 4 * (k >> 1) + (k & 1);
      const uint32_t m0 = max(top[i0], top[i0 + 2]);  // pooled row 0 of the block, column j
      const uint32_t m1 = max(bot[i0], bot[i0 + 2]);  // pooled row 1
      const uint32_t send = odd ? m0 : m1;
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 4);
      r[k] = max(odd ? m1 : m0, recv);
    }
    uint32_t bad = 0;
    if (fast) {
      if (u == 0) packed[0] = epi16_requant4<CHECK, 0>(r[0], r[1], r[2], r[3], K, zp_sub, lo, bad);
      if (u == 1) packed[1] = epi16_requant4<CHECK, 4>(r[0], r[1], r[2], r[3], K, zp_sub, lo, bad);
      if (u == 2) packed[2] = epi16_requant4<CHECK, 8>(r[0], r[1], r[2], r[3], K, zp_sub, lo, bad);
      if (u == 3) packed[3] = epi16_requant4<CHECK, 12>(r[0], r[1], r[2], r[3], K, zp_sub, lo, bad);
    }
    if (!fast || (CHECK && __any_sync(0xffffffffu, requant_magic_out_of_range(bad))))
      packed[u] = epi16_requant4_exact(r[0], r[1], r[2], r[3], consts, ch0 + 4 * u, zp_out, lo);
  }
  if (valid) *reinterpret_cast<uint4*>(out) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
}

}  // namespace b200q
