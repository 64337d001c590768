// tcgen05 int8 implicit-GEMM: 3x3/s1/p1 quantized convolution and quantized linear for sm_100a.
//
//   D[M=128 pixels][N=cout] (s32, TMEM)  +=  A[128][K chunk] (u8 activations, smem)  x  B[N][K chunk]^T (s8 weights, smem)
//
// * A tiles are im2col views fetched by TMA straight from the NHWC activation tensor: one 4-D box
//   {KC channels, IMG columns, rows, images} per filter tap, shifted by (kw-1, kh-1); out-of-image elements are
//   zero-filled by TMA.  Quantized zero (zp_x) is NOT 0, so the epilogue subtracts zp_x * sum(valid-tap weights),
//   looked up per border class (3 row classes x 3 column classes) and output channel.
// * B tiles (weights [cout][9*cin], K-major) are either streamed with A or, when the whole layer fits, loaded once
//   and kept resident in shared memory for the lifetime of the persistent CTA.
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 2..5 = epilogue (one TMEM lane
//   quarter each).  Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// * Epilogue: tcgen05.ld -> exact fbgemm requantisation (common.cuh) -> packed uint8 NHWC stores.
#include "common.cuh"

namespace b200q {

constexpr int TC_THREADS = 192;
constexpr int TILE_M = 128;
constexpr int SMEM_BUDGET = 220 * 1024;

template <int IMG, int CIN, int COUT, bool B_RESIDENT>
struct TcCfg {
  static constexpr bool CONV = IMG > 0;
  static constexpr int KC = (CIN % 128 == 0) ? 128 : 64;  // K-chunk bytes == swizzle span
  static constexpr int TAPS = CONV ? 9 : 1;
  static constexpr int CHUNKS_PER_TAP = CIN / KC;
  static constexpr int NCHUNK = TAPS * CHUNKS_PER_TAP;
  static constexpr int N_TILE = COUT > 256 ? 256 : COUT;
  static constexpr int N_TILES = COUT / N_TILE;
  static constexpr int A_BYTES = TILE_M * KC;
  static constexpr int B_BYTES = N_TILE * KC;
  static constexpr int CFGS = CONV ? 9 : 1;
  static constexpr int TABLE_BYTES = (CFGS + 2) * COUT * 4;
  static constexpr int B_SLOTS_RESIDENT = NCHUNK;
  static constexpr int FIXED = TABLE_BYTES + 1024 /*barriers etc*/ + 1024 /*alignment slack*/;
  static constexpr int AVAIL = SMEM_BUDGET - FIXED - (B_RESIDENT ? B_SLOTS_RESIDENT * B_BYTES : 0);
  static constexpr int STAGE_BYTES = A_BYTES + (B_RESIDENT ? 0 : B_BYTES);
  static constexpr int STAGES_RAW = AVAIL / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int B_SLOTS = B_RESIDENT ? B_SLOTS_RESIDENT : STAGES;
  static constexpr int SMEM_BYTES = STAGES * A_BYTES + B_SLOTS * B_BYTES + FIXED;
  static constexpr int TMEM_COLS = 2 * N_TILE;  // 128 / 256 / 512: power of two >= 32
  // conv tile geometry: 128 pixels = NB images x ROWS rows x IMG columns
  static constexpr int ROWS = CONV ? (TILE_M / IMG > IMG ? IMG : TILE_M / IMG) : 1;
  static constexpr int NB = CONV ? TILE_M / (ROWS * IMG) : 1;
  static constexpr int TILES_PER_IMG = CONV ? (IMG / ROWS) : 1;  // 8, 2, 1(=covers NB images)
  static_assert(CIN % KC == 0 && COUT % N_TILE == 0, "shape");
  static_assert(N_TILE % 32 == 0 && N_TILE >= 32 && N_TILE <= 256, "N tile");
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
  static_assert(!B_RESIDENT || N_TILES == 1, "resident weights need a single N tile");
  static_assert(!CONV || ROWS * NB * IMG == TILE_M, "tile geometry");
};

struct TcArgs {
  uint8_t* y;
  const float* mult;
  const float* bdiv;
  const int32_t* corr;
  int64_t m_rows;      // conv: number of images; linear: number of rows
  int num_m_tiles;
  int zp_out, lo;
};

template <int IMG, int CIN, int COUT, bool B_RESIDENT>
__global__ void __launch_bounds__(TC_THREADS, 1)
igemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const TcArgs args) {
  using C = TcCfg<IMG, CIN, COUT, B_RESIDENT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;
  uint8_t* b_smem = a_smem + C::STAGES * C::A_BYTES;
  int32_t* s_corr = reinterpret_cast<int32_t*>(b_smem + C::B_SLOTS * C::B_BYTES);
  float* s_mult = reinterpret_cast<float*>(s_corr + C::CFGS * COUT);
  float* s_bdiv = s_mult + COUT;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_bdiv + COUT);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = args.num_m_tiles * C::N_TILES;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full_bar + i, 1);
      mbar_init(tmem_empty_bar + i, 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < C::CFGS * COUT; i += 128) s_corr[i] = __ldg(args.corr + i);
    for (int i = threadIdx.x - 64; i < COUT; i += 128) {
      s_mult[i] = __ldg(args.mult + i);
      s_bdiv[i] = __ldg(args.bdiv + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      bool first_tile = true;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / C::N_TILES;
        const int n_tile = tile % C::N_TILES;
        for (int j = 0; j < C::NCHUNK; ++j) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          const bool load_b = !B_RESIDENT || first_tile;
          mbar_expect_tx(full_bar + stage, C::A_BYTES + (load_b ? C::B_BYTES : 0));
          uint8_t* a_dst = a_smem + stage * C::A_BYTES;
          if constexpr (C::CONV) {
            const int tap = j / C::CHUNKS_PER_TAP;
            const int c0 = (j % C::CHUNKS_PER_TAP) * C::KC;
            const int kh = tap / 3, kw = tap % 3;
            const int img0 = (m_tile / C::TILES_PER_IMG) * C::NB;
            const int row0 = (m_tile % C::TILES_PER_IMG) * C::ROWS;
            tma_load_4d(a_dst, &map_a, full_bar + stage, c0, kw - 1, row0 + kh - 1, img0);
          } else {
            tma_load_2d(a_dst, &map_a, full_bar + stage, j * C::KC, m_tile * TILE_M);
          }
          if (load_b) {
            uint8_t* b_dst = b_smem + (B_RESIDENT ? j : (int)stage) * C::B_BYTES;
            tma_load_2d(b_dst, &map_b, full_bar + stage, j * C::KC, n_tile * C::N_TILE);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        first_tile = false;
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_i8(TILE_M, C::N_TILE);
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t slot = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(tmem_empty_bar + slot, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + slot * C::N_TILE;
        for (int j = 0; j < C::NCHUNK; ++j) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_smem + stage * C::A_BYTES);
          const uint32_t b_addr = smem_u32(b_smem + (B_RESIDENT ? j : (int)stage) * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < C::KC / 32; ++k) {
            const uint64_t da = make_kmajor_desc<C::KC>(a_addr + k * 32, 8 * C::KC);
            const uint64_t db = make_kmajor_desc<C::KC>(b_addr + k * 32, 8 * C::KC);
            tc_mma_i8(d_tmem, da, db, idesc, (j | k) != 0 ? 1u : 0u);
          }
          tc_commit(empty_bar + stage);  // smem slot reusable once these MMAs have read it
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(tmem_full_bar + slot);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ================================================================== epilogue warps (TMEM lane quarter = warp % 4)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_tile = tile / C::N_TILES;
      const int n0 = (tile % C::N_TILES) * C::N_TILE;
      const uint32_t slot = it & 1, acc_phase = (it >> 1) & 1;

      int cfg = 0;
      bool valid;
      uint8_t* out;
      if constexpr (C::CONV) {
        const int64_t img = (int64_t)(m_tile / C::TILES_PER_IMG) * C::NB + row / (C::ROWS * IMG);
        const int h = (m_tile % C::TILES_PER_IMG) * C::ROWS + (row / IMG) % C::ROWS;
        const int w = row % IMG;
        cfg = (h == 0 ? 0 : (h == IMG - 1 ? 2 : 1)) * 3 + (w == 0 ? 0 : (w == IMG - 1 ? 2 : 1));
        valid = img < args.m_rows;
        out = args.y + ((img * IMG + h) * IMG + w) * (int64_t)COUT + n0;
      } else {
        const int64_t r = (int64_t)m_tile * TILE_M + row;
        valid = r < args.m_rows;
        out = args.y + r * (int64_t)COUT + n0;
      }
      const int32_t* corr_row = s_corr + cfg * COUT + n0;

      mbar_wait(tmem_full_bar + slot, acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * C::N_TILE;
#pragma unroll 1
      for (int c0 = 0; c0 < C::N_TILE; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_addr + c0, v);
        tmem_ld_wait();
        uint32_t packed[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint32_t word = 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c0 + g * 4 + e;
            const int acc = (int)v[g * 4 + e] - corr_row[c];
            word |= requant_u8(acc, s_bdiv[n0 + c], s_mult[n0 + c], args.zp_out, args.lo) << (8 * e);
          }
          packed[g] = word;
        }
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(out + c0);
          dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar + slot);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// --------------------------------------------------------------------------------------------------------- host side
template <int IMG, int CIN, int COUT, bool B_RESIDENT>
static int launch_tc(const uint8_t* x, uint8_t* y, int64_t m_rows, const int8_t* w, const int32_t* corr,
                     const b200q_requant& rq, cudaStream_t stream) {
  using C = TcCfg<IMG, CIN, COUT, B_RESIDENT>;
  CUtensorMap map_a, map_b;
  int num_m_tiles;
  int rc;
  if constexpr (C::CONV) {
    const uint64_t dims[4] = {(uint64_t)CIN, (uint64_t)IMG, (uint64_t)IMG, (uint64_t)m_rows};
    const uint64_t strides[3] = {(uint64_t)CIN, (uint64_t)IMG * CIN, (uint64_t)IMG * IMG * CIN};
    const uint32_t box[4] = {(uint32_t)C::KC, (uint32_t)IMG, (uint32_t)C::ROWS, (uint32_t)C::NB};
    rc = encode_tensor_map(&map_a, x, 4, dims, strides, box, C::KC);
    if (rc) return rc;
    num_m_tiles = (int)(((m_rows + C::NB - 1) / C::NB) * C::TILES_PER_IMG);
  } else {
    const uint64_t dims[2] = {(uint64_t)CIN, (uint64_t)m_rows};
    const uint64_t strides[1] = {(uint64_t)CIN};
    const uint32_t box[2] = {(uint32_t)C::KC, (uint32_t)TILE_M};
    rc = encode_tensor_map(&map_a, x, 2, dims, strides, box, C::KC);
    if (rc) return rc;
    num_m_tiles = (int)((m_rows + TILE_M - 1) / TILE_M);
  }
  {
    const uint64_t ktot = (uint64_t)C::TAPS * CIN;
    const uint64_t dims[2] = {ktot, (uint64_t)COUT};
    const uint64_t strides[1] = {ktot};
    const uint32_t box[2] = {(uint32_t)C::KC, (uint32_t)C::N_TILE};
    rc = encode_tensor_map(&map_b, w, 2, dims, strides, box, C::KC);
    if (rc) return rc;
  }
  auto kernel = igemm_tc_kernel<IMG, CIN, COUT, B_RESIDENT>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    B200Q_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  TcArgs args{y, rq.mult, rq.bdiv, corr, m_rows, num_m_tiles, rq.zp_out, rq.relu ? rq.zp_out : 0};
  const int num_tiles = num_m_tiles * C::N_TILES;
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  kernel<<<grid, TC_THREADS, C::SMEM_BYTES, stream>>>(map_a, map_b, args);
  return launched("igemm_tc_kernel");
}

}  // namespace b200q

using namespace b200q;

// B200Q_TC_STREAM_WEIGHTS=1 forces the streamed-weights variant for layers that default to resident weights
// (bring-up / A-B testing only).
static bool force_streamed() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200Q_TC_STREAM_WEIGHTS");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

extern "C" int b200q_conv3x3_tc(const uint8_t* x, uint8_t* y, int64_t b, const b200q_conv3x3* L, int pool2x2,
                                void* stream) {
  B200Q_REQUIRE(L && ((x && y) || b == 0), "conv3x3_tc: null pointer");
  B200Q_REQUIRE(L->w && L->corr && L->rq.mult && L->rq.bdiv, "conv3x3_tc: unpacked layer");
  B200Q_REQUIRE(pool2x2 == 0, "conv3x3_tc: fused 2x2 max-pool not available in this build");
  B200Q_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)L->w % 16 == 0,
                "conv3x3_tc: buffers must be 16-byte aligned");
  if (b == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const bool streamed = force_streamed();
#define B200Q_TC_CASE(IMG_, CIN_, COUT_, RES_)                                                       \
  if (L->img == IMG_ && L->cin == CIN_ && L->cout == COUT_) {                                        \
    if (RES_ && !streamed) return launch_tc<IMG_, CIN_, COUT_, RES_>(x, y, b, L->w, L->corr, L->rq, s); \
    return launch_tc<IMG_, CIN_, COUT_, false>(x, y, b, L->w, L->corr, L->rq, s);                    \
  }
  B200Q_TC_CASE(32, 64, 64, true)
  B200Q_TC_CASE(16, 64, 128, true)
  B200Q_TC_CASE(16, 128, 128, true)
  B200Q_TC_CASE(8, 128, 256, false)
  B200Q_TC_CASE(8, 256, 256, false)
#undef B200Q_TC_CASE
  set_error("conv3x3_tc: unsupported geometry img=%d cin=%d cout=%d", L->img, L->cin, L->cout);
  return B200Q_ERR_INVALID_ARG;
}

extern "C" int b200q_linear_tc(const uint8_t* x, uint8_t* y, int64_t b, const b200q_linear* L, void* stream) {
  B200Q_REQUIRE(L && ((x && y) || b == 0), "linear_tc: null pointer");
  B200Q_REQUIRE(L->w && L->corr && L->rq.mult && L->rq.bdiv, "linear_tc: unpacked layer");
  B200Q_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)L->w % 16 == 0,
                "linear_tc: buffers must be 16-byte aligned");
  if (b == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (L->k == 4096 && L->n == 512) return launch_tc<0, 4096, 512, false>(x, y, b, L->w, L->corr, L->rq, s);
  if (L->k == 512 && L->n == 64) return launch_tc<0, 512, 64, false>(x, y, b, L->w, L->corr, L->rq, s);
  set_error("linear_tc: unsupported geometry k=%d n=%d", L->k, L->n);
  return B200Q_ERR_INVALID_ARG;
}
