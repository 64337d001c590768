// quantized::linear_dynamic(x, W, reduce_range=True) on the tensor cores (nnqd.Linear as quantize_dynamic builds it,
// models/dynamic_ptq_model.py:302-306): after ONE min/max pass over the input (elementwise.cu minmax_kernel, which also
// derives fbgemm's (scale, zero_point) on the device) a single kernel does everything else:
//
//   fp32 x [b][k]  --producer warps: rne(fma(x, 1/s_x, zp)) -> saturate to u8----->  swizzled K-major smem tiles
//   int8 W [n][k]  --TMA----------------------------------------------------------->  swizzled K-major smem tiles
//   tcgen05.mma.kind::i8 (M=128, N<=256, K=32), s32 accumulators in TMEM
//   epilogue: y = f32(acc - zp * wsum[n]) * (s_x * s_w) + bias[n]  (+ ReLU), fp32 [b][n]
//
// The quantised activations never exist in HBM: the layer reads its fp32 input exactly twice (min/max pass + this
// kernel) and is bound by that stream (fc1: 16 KB per image against 4.2 MOP).  A CTA owns 128 rows and ALL n output
// columns (n = 512: two N=256 accumulators = the whole TMEM), so x is quantised once per row block; the weights are
// re-streamed per row block from L2 (2 MB for fc1; 64-byte K chunks).
//
// The fp32 rows are staged by TMA: a [128 rows][64 floats] box (32 KB) per K chunk into a three-deep shared-memory ring,
// i.e. 96 KB of HBM reads in flight per SM without a single register (the first version loaded through registers with one
// chunk of look-ahead and ran at 2.85 TB/s: every chunk exposed an HBM round trip; a deeper register ring was slower
// still - ptxas folded the in-flight loads onto shared scoreboards).  Rows past the batch are zero-filled by TMA.
//
// Warp roles (448 threads): warps 0..3 epilogue (warp q may only read TMEM lanes 32q..32q+31), warps 4..11 producers
// (shared fp32 -> shared u8), warp 12 TMA (fp32 rows and weights), warp 13 MMA issuer + TMEM owner.
#include "common.cuh"

namespace b200q {

int launch_minmax(const float* x, int64_t n, float* out5, void* scratch, cudaStream_t s, bool dynamic);

constexpr int LD_EPI_WARPS = 4, LD_PROD_WARPS = 8;
constexpr int LD_PROD_WARP0 = LD_EPI_WARPS;
constexpr int LD_TMA_WARP = LD_EPI_WARPS + LD_PROD_WARPS, LD_MMA_WARP = LD_TMA_WARP + 1;
constexpr int LD_THREADS = 32 * (LD_MMA_WARP + 1);
constexpr int LD_M = 128;
constexpr int LD_KC = 64;  // K bytes per stage == swizzle span (SWIZZLE_64B)

template <int NT>
struct LdCfg {
  static constexpr int MMA_N = NT > 256 ? 256 : NT;
  static constexpr int N_MMAS = NT / MMA_N;
  static constexpr int A_BYTES = LD_M * LD_KC, B_BYTES = NT * LD_KC;
  static constexpr int RAW_BYTES = LD_M * LD_KC * 4;  // one K chunk of fp32 rows
  static constexpr int RAW_STAGES = 3;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_FIT = (216 * 1024 - RAW_STAGES * RAW_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr int SMEM_BYTES = RAW_STAGES * RAW_BYTES + STAGES * STAGE_BYTES + 256 + 1024;
  static constexpr int TMEM_COLS = NT < 32 ? 32 : NT;
  static_assert(NT % MMA_N == 0 && MMA_N % 16 == 0 && MMA_N >= 16, "N tile");
  static_assert(B_BYTES % 1024 == 0, "stage alignment");
  static_assert(STAGES >= 2, "pipeline");
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns");
};

struct LdArgs {
  const float* x;
  float* y;
  const int32_t* wsum;
  const float* bias;
  const float* qp;   // device {min, max, scale, 1/scale, zp}
  int64_t b;
  int k, n;
  int num_m_tiles;
  int relu;
  float w_scale;
};

template <int NT>
__global__ void __launch_bounds__(LD_THREADS, 1)
linear_dynamic_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                         const LdArgs args) {
  using C = LdCfg<NT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                                   // [STAGES][128][64 B]
  uint8_t* b_smem = a_smem + C::STAGES * C::A_BYTES;        // [STAGES][NT][64 B]
  uint8_t* x_smem = b_smem + C::STAGES * C::B_BYTES;        // [RAW_STAGES][128][64 floats]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(x_smem + C::RAW_STAGES * C::RAW_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* raw_full_bar = empty_bar + C::STAGES;           // [RAW_STAGES] fp32 chunk landed (TMA)
  uint64_t* raw_empty_bar = raw_full_bar + C::RAW_STAGES;   // [RAW_STAGES] fp32 chunk converted by all producer warps
  uint64_t* tmem_full_bar = raw_empty_bar + C::RAW_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 1;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nk = args.k / LD_KC;

  if (warp == LD_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(full_bar + i, LD_PROD_WARPS + 1);  // one arrive per producer warp + the TMA thread's arrive.expect_tx
      mbar_init(empty_bar + i, 1);
    }
    for (int i = 0; i < C::RAW_STAGES; ++i) {
      mbar_init(raw_full_bar + i, 1);
      mbar_init(raw_empty_bar + i, LD_PROD_WARPS);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, LD_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == LD_MMA_WARP) {
    tmem_alloc(tmem_base_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp >= LD_PROD_WARP0 && warp < LD_TMA_WARP) {
    // ================================================================== producers: shared fp32 rows -> u8 SW64 tiles
    // A warp-wide 16-byte shared load covers two rows of the chunk (2 x 64 floats, 512 contiguous bytes: conflict-free);
    // warp pw owns rows 16*pw .. 16*pw+15, thread (lane) converts floats 4*(lane&15)..+3 of rows 16*pw + 2*i + (lane>>4).
    const int pw = warp - LD_PROD_WARP0;
    const float inv_scale = __ldg(args.qp + 3);
    const float zp_f = __ldg(args.qp + 4);
    const int col4 = lane & 15, rsub = lane >> 4;
    // fbgemm quantises the activations of a dynamic linear with the zero-point added in fp32 BEFORE the rounding to
    // integer, as ONE fused multiply-add: q = clamp(rne(fma(x, 1/s, zp)), 0, 255) (PackAWithQuantRowOffset; found by
    // experiment - 0 mismatches over 4.2e7 elements against quantized::linear_dynamic, while rne(x/s)+zp misses 134 and
    // rne(fl(x/s)+zp) 33 of them; oracle/int_ops.linear_dynamic, tests/test_oracle.py).  |fma(...)| <= ~255 by
    // construction of the scale, so the round-to-nearest-even of the fp32 adder itself (t + 1.5*2^23) is exact and the
    // conversion pipe is not needed; the saturating pack clamps both ends.
    constexpr int UNMAGIC = -(int)MAGIC_BITS;
    auto quant4 = [&](const float4 v) -> uint32_t {
      const int q0 = __float_as_int(__fadd_rn(__fmaf_rn(v.x, inv_scale, zp_f), MAGIC_F)) + UNMAGIC;
      const int q1 = __float_as_int(__fadd_rn(__fmaf_rn(v.y, inv_scale, zp_f), MAGIC_F)) + UNMAGIC;
      const int q2 = __float_as_int(__fadd_rn(__fmaf_rn(v.z, inv_scale, zp_f), MAGIC_F)) + UNMAGIC;
      const int q3 = __float_as_int(__fadd_rn(__fmaf_rn(v.w, inv_scale, zp_f), MAGIC_F)) + UNMAGIC;
      return pack_sat_u8(q1, q0, pack_sat_u8(q3, q2, 0u));
    };
    uint32_t stage = 0, phase = 0, rstage = 0, rphase = 0;
    for (int tile = blockIdx.x; tile < args.num_m_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(raw_full_bar + rstage, rphase);
        const float4* src = reinterpret_cast<const float4*>(x_smem + rstage * C::RAW_BYTES);
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = src[(16 * pw + 2 * i + rsub) * (LD_KC / 4) + col4];
        mbar_wait(empty_bar + stage, phase ^ 1);
        uint8_t* a_st = a_smem + stage * C::A_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 16 * pw + 2 * i + rsub;
          // SWIZZLE_64B: 16-byte chunk index ^= address bits [7,9) = (row >> 1) & 3
          const int off = r * LD_KC + ((((col4 >> 2) ^ ((r >> 1) & 3))) << 4) + (col4 & 3) * 4;
          *reinterpret_cast<uint32_t*>(a_st + off) = quant4(v[i]);
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(raw_empty_bar + rstage);  // this warp's rows of the fp32 chunk are consumed
          mbar_arrive(full_bar + stage);
        }
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
        if (++rstage == C::RAW_STAGES) {
          rstage = 0;
          rphase ^= 1;
        }
      }
    }
  } else if (warp == LD_TMA_WARP) {
    // ================================================================== fp32 rows and weights by TMA
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, rstage = 0, rphase = 0;
      for (int tile = blockIdx.x; tile < args.num_m_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(raw_empty_bar + rstage, rphase ^ 1);
          mbar_expect_tx(raw_full_bar + rstage, C::RAW_BYTES);  // rows past the batch are zero-filled, still counted
          tma_load_2d(x_smem + rstage * C::RAW_BYTES, &map_x, raw_full_bar + rstage, kc * LD_KC * 4, tile * LD_M);
          if (++rstage == C::RAW_STAGES) {
            rstage = 0;
            rphase ^= 1;
          }
          mbar_wait(empty_bar + stage, phase ^ 1);
          mbar_expect_tx(full_bar + stage, C::B_BYTES);
          uint8_t* b_st = b_smem + stage * C::B_BYTES;
#pragma unroll
          for (int h = 0; h < C::N_MMAS; ++h)
            tma_load_2d(b_st + h * C::MMA_N * LD_KC, &map_w, full_bar + stage, kc * LD_KC, h * C::MMA_N);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == LD_MMA_WARP) {
    // ================================================================== MMA issuer
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc_i8(LD_M, C::MMA_N);
    const uint64_t a_desc0 = make_kmajor_desc<LD_KC>(smem_u32(a_smem), 8 * LD_KC);
    const uint64_t b_desc0 = make_kmajor_desc<LD_KC>(smem_u32(b_smem), 8 * LD_KC);
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < args.num_m_tiles; tile += gridDim.x, ++it) {
      mbar_wait(tmem_empty_bar, (it & 1) ^ 1);  // single accumulator set: the epilogue of the previous tile has drained it
      tc_fence_after();
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(full_bar + stage, phase);
        tc_fence_after();
        if (leader) {
          const uint64_t da0 = a_desc0 + (uint64_t)((stage * C::A_BYTES) >> 4);
          const uint64_t db0 = b_desc0 + (uint64_t)((stage * C::B_BYTES) >> 4);
#pragma unroll
          for (int ks = 0; ks < LD_KC / 32; ++ks)
#pragma unroll
            for (int h = 0; h < C::N_MMAS; ++h)
              tc_mma_i8(tmem_base + h * C::MMA_N, da0 + (uint64_t)((ks * 32) >> 4),
                        db0 + (uint64_t)((h * C::MMA_N * LD_KC + ks * 32) >> 4), idesc, (kc | ks) != 0 ? 1u : 0u);
          tc_commit(empty_bar + stage);
          if (kc == nk - 1) tc_commit(tmem_full_bar);
        }
        __syncwarp();
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ================================================================== epilogue: fp32 out
    const int quarter = warp;
    const float s_x = __ldg(args.qp + 2);
    const int zp = (int)__ldg(args.qp + 4);
    const float s_xw = __fmul_rn(s_x, args.w_scale);
    constexpr int CW = NT >= 32 ? 32 : 16;  // columns per tcgen05.ld
    int it = 0;
    for (int tile = blockIdx.x; tile < args.num_m_tiles; tile += gridDim.x, ++it) {
      const int64_t row = (int64_t)tile * LD_M + quarter * 32 + lane;
      const bool valid = row < args.b;
      mbar_wait(tmem_full_bar, it & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += CW) {
        uint32_t v[32];
        if constexpr (CW == 32) {
          tmem_ld_32x32(t_addr + c0, v);
        } else {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(t_addr + c0)
              : "memory");
        }
        tmem_ld_wait();
        if (c0 + CW >= NT) {  // accumulators read: hand TMEM back before the arithmetic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar);
        }
        float o[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) {
          const int nn = c0 + j;
          const int nc = nn < args.n ? nn : args.n - 1;  // padded columns (n < NT): computed, never stored
          const int t = (int)v[j] - zp * __ldg(args.wsum + nc);
          // fbgemm's output stage is ONE fused multiply-add, fma(f32(acc), s_x*s_w, bias), with the scale product rounded
          // to fp32 first (found by experiment: bit-identical to quantized::linear_dynamic on every element tried,
          // tests/test_oracle.py; separately rounded mul + add differs in the last place on ~25 % of them)
          float r = __fmaf_rn(__int2float_rn(t), s_xw, __ldg(args.bias + nc));
          o[j] = args.relu ? fmaxf(r, 0.0f) : r;
        }
        if (valid) {
          float* dst = args.y + row * args.n + c0;
          if (args.n % 4 == 0 && c0 + CW <= args.n) {
#pragma unroll
            for (int j = 0; j < CW; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < CW; ++j)
              if (c0 + j < args.n) dst[j] = o[j];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == LD_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int NT>
static int launch_linear_dynamic(const LdArgs& a, const int8_t* w, cudaStream_t s) {
  using C = LdCfg<NT>;
  CUtensorMap map_x, map_w;
  {  // fp32 rows as bytes: [b][4k] uint8, box = 256 bytes (64 floats) x 128 rows, no swizzle
    const uint64_t xdims[2] = {(uint64_t)a.k * 4, (uint64_t)a.b};
    const uint64_t xstrides[1] = {(uint64_t)a.k * 4};
    const uint32_t xbox[2] = {(uint32_t)LD_KC * 4, (uint32_t)LD_M};
    if (int rc = encode_tensor_map(&map_x, a.x, 2, xdims, xstrides, xbox, 0)) return rc;
  }
  const uint64_t dims[2] = {(uint64_t)a.k, (uint64_t)a.n};
  const uint64_t strides[1] = {(uint64_t)a.k};
  const uint32_t box[2] = {(uint32_t)LD_KC, (uint32_t)C::MMA_N};  // rows >= n (n < NT) are zero-filled by TMA
  if (int rc = encode_tensor_map(&map_w, w, 2, dims, strides, box, LD_KC)) return rc;
  auto kernel = linear_dynamic_tc_kernel<NT>;
  static uint64_t attr_mask = 0;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), C::SMEM_BYTES, &attr_mask)) return rc;
  const int grid = a.num_m_tiles < num_sms() ? a.num_m_tiles : num_sms();
  kernel<<<grid, LD_THREADS, C::SMEM_BYTES, s>>>(map_x, map_w, a);
  return launched("linear_dynamic_tc_kernel");
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_linear_dynamic(const float* x, float* y, int64_t b, int k, int n, const int8_t* w,
                                    const int32_t* wsum, float w_scale, const float* bias, int relu, void* scratch,
                                    int64_t scratch_bytes, void* stream) {
  B200Q_REQUIRE(b >= 0, "linear_dynamic: negative batch");
  if (b == 0) return 0;
  B200Q_REQUIRE(x && y && w && wsum && bias && scratch, "linear_dynamic: null pointer");
  B200Q_REQUIRE(scratch_bytes >= B200Q_REDUCE_SCRATCH_BYTES, "linear_dynamic: scratch too small (%lld < %d)",
                (long long)scratch_bytes, B200Q_REDUCE_SCRATCH_BYTES);
  B200Q_REQUIRE(k > 0 && k % 64 == 0 && (n == 512 || (n > 0 && n <= 16)),
                "linear_dynamic: unsupported shape k=%d n=%d (need k %% 64 == 0 and n == 512 or n <= 16)", k, n);
  B200Q_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)w % 16 == 0 && (uintptr_t)scratch % 16 == 0,
                "linear_dynamic: buffers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  float* qp = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + B200Q_REDUCE_QPARAMS_OFFSET);
  if (int rc = launch_minmax(x, b * k, qp, scratch, s, true)) return rc;
  LdArgs a{x, y, wsum, bias, qp, b, k, n, (int)((b + LD_M - 1) / LD_M), relu, w_scale};
  return n == 512 ? launch_linear_dynamic<512>(a, w, s) : launch_linear_dynamic<16>(a, w, s);
}
