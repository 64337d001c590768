// tcgen05 int8 implicit-GEMM: 3x3/s1/p1 quantized convolution (+ fused 2x2 max-pool) and quantized linear for sm_100a.
//
//   D[M=128 pixels][N=cout] (s32, TMEM)  +=  A[128][K chunk] (u8 activations, smem)  x  B[N][K chunk]^T (s8 weights, smem)
//
// * A tiles are im2col views fetched by TMA straight from the NHWC activation tensor: one 4-D box
//   {KC channels, IMG columns, rows, images} per filter tap, shifted by (kw-1, kh-1); out-of-image elements are
//   zero-filled by TMA.  Quantized zero (zp_x) is NOT 0, so the epilogue subtracts zp_x * sum(valid-tap weights),
//   looked up per border class (3 row classes x 3 column classes) and output channel.
// * B tiles (weights [cout][9*cin], K-major) are either streamed with A or, when the whole layer fits, loaded once
//   and kept resident in shared memory for the lifetime of the persistent CTA.
// * Warp roles (320 threads): warps 0..7 = epilogue (two per TMEM lane quarter, each taking half of the N columns),
//   warp 8 = TMA producer, warp 9 = MMA issuer (+TMEM owner; highest warp id = highest arbitration priority).  Accumulators are double-buffered in TMEM so the
//   epilogue of tile i overlaps the MMAs of tile i+1; a warp releases its accumulator slot right after its last
//   tcgen05.ld, before doing the arithmetic.
// * Epilogue: tcgen05.ld -> exact fbgemm requantisation -> packed uint8.  The arithmetic avoids the conversion pipe
//   (common.cuh: requant4_magic) and falls back to the I2F/F2I form when a range check fails or the layer's constants
//   are not flagged B200Q_RQ_BOUNDED.  Without pooling each thread stores its pixel's 32-channel chunk directly (NHWC);
//   with pooling the tile is staged in swizzled shared memory, 2x2-max-reduced (max commutes with the monotone
//   requantisation, so this equals aten::quantized_max_pool2d on the stored tensor) and written pooled.
#include "common.cuh"

namespace b200q {

constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_TMA_WARP = TC_EPI_WARPS, TC_MMA_WARP = TC_EPI_WARPS + 1;
constexpr int TILE_M = 128;
constexpr int SMEM_BUDGET = 222 * 1024;

// NT: N tile (accumulator columns per CTA tile); 0 = the widest that fits (min(COUT, 256)).  NT = 64 is the SMALL-BATCH
// shape: a layer is cut into COUT/64 times more CTA tiles, so that a handful of images still spreads over many SMs (the
// whole-network executor, net.cu b200q_graph_*: at batch 1 every layer is a chain of dependent L2 round trips, and the
// only lever is how little each CTA has to stream).
template <int IMG, int CIN, int COUT, bool B_RESIDENT, bool POOL, int NT = 0>
struct TcCfg {
  static constexpr bool CONV = IMG > 0;
  static constexpr int KC = (CIN % 128 == 0) ? 128 : 64;  // K-chunk bytes == swizzle span
  static constexpr int TAPS = CONV ? 9 : 1;
  static constexpr int CHUNKS_PER_TAP = CIN / KC;
  static constexpr int NCHUNK = TAPS * CHUNKS_PER_TAP;
  static constexpr int N_TILE = NT > 0 ? NT : (COUT > 256 ? 256 : COUT);
  static constexpr int N_TILES = COUT / N_TILE;
  static constexpr int A_BYTES = TILE_M * KC;
  static constexpr int B_BYTES = N_TILE * KC;
  static constexpr int CFGS = CONV ? 9 : 1;
  static constexpr int CM_STRIDE = COUT + 4;  // words; +16 B so the border classes of one warp hit different banks
  static constexpr int TABLE_BYTES = (CFGS * CM_STRIDE + 2 * COUT) * 4;
  static constexpr int STAGING_BYTES = POOL ? TILE_M * N_TILE : 0;
  static constexpr int B_SLOTS_RESIDENT = NCHUNK;
  static constexpr int FIXED = TABLE_BYTES + STAGING_BYTES + 1024 /*barriers etc*/ + 1024 /*alignment slack*/;
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
  // epilogue split: each warp owns one TMEM lane quarter and N_TILE/2 columns
  static constexpr int COLS_PER_WARP = N_TILE / 2;
  static constexpr int CHUNKS_PER_WARP = COLS_PER_WARP / 32;
  static_assert(CIN % KC == 0 && COUT % N_TILE == 0, "shape");
  static_assert(N_TILE % 64 == 0 && N_TILE >= 64 && N_TILE <= 256, "N tile");
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
  static_assert(!B_RESIDENT || N_TILES == 1, "resident weights need a single N tile");
  static_assert(!CONV || ROWS * NB * IMG == TILE_M, "tile geometry");
  static_assert(!POOL || (CONV && ROWS % 2 == 0), "pool fusion needs whole 2x2 windows in a tile");
};

struct TcArgs {
  uint8_t* y;
  const float* mult;
  const float* bdiv;
  const int32_t* corr;
  int64_t m_rows;      // conv: number of images; linear: number of rows
  int num_m_tiles;
  int zp_out, lo;
  int bounded;         // B200Q_RQ_BOUNDED: the conversion-free requantisation is exact for in-range accumulators
};

// Staging layout for the pooled epilogue: row p (pixel of the tile) holds N bytes; its 16-byte chunk j lives at
// chunk slot j ^ swz(p) so that both the per-pixel writes and the per-window reads spread over the banks.
template <int N>
__device__ __forceinline__ int staging_off(int p, int j) {
  const int f = (N == 64) ? ((p >> 1) & 3) : (p & 7);
  return p * N + ((j ^ f) << 4);
}

template <int IMG, int CIN, int COUT, bool B_RESIDENT, bool POOL, bool CHECK, int NT>
__global__ void __launch_bounds__(TC_THREADS, 1)
igemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const TcArgs args) {
  using C = TcCfg<IMG, CIN, COUT, B_RESIDENT, POOL, NT>;
  extern __shared__ uint8_t smem_raw[];
  // 1 KiB alignment for the 128B-swizzle atoms; plain pointer arithmetic keeps the shared address space visible to
  // the compiler (LDS/STS instead of generic loads)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem = smem;
  uint8_t* b_smem = a_smem + C::STAGES * C::A_BYTES;
  uint8_t* staging = b_smem + C::B_SLOTS * C::B_BYTES;  // 1024-aligned (all tile sizes are multiples of 1 KiB)
  int32_t* s_cm = reinterpret_cast<int32_t*>(staging + C::STAGING_BYTES);
  float* s_mult = reinterpret_cast<float*>(s_cm + C::CFGS * C::CM_STRIDE);
  float* s_bdiv = s_mult + COUT;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_bdiv + COUT);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = args.num_m_tiles * C::N_TILES;
  pdl_launch_dependents();

  if (warp == TC_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full_bar + i, 1);
      mbar_init(tmem_empty_bar + i, TC_EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == TC_MMA_WARP) {
    tmem_alloc(tmem_base_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp < TC_EPI_WARPS) {
    const int t = threadIdx.x;
    for (int i = t; i < C::CFGS * COUT; i += 32 * TC_EPI_WARPS)
      s_cm[(i / COUT) * C::CM_STRIDE + i % COUT] = (int32_t)(MAGIC_BITS - (uint32_t)__ldg(args.corr + i));
    for (int i = t; i < COUT; i += 32 * TC_EPI_WARPS) {
      s_mult[i] = __ldg(args.mult + i);
      s_bdiv[i] = __ldg(args.bdiv + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == TC_TMA_WARP) {
    // ================================================================== TMA producer
    if (lane == 0) {
      pdl_wait();  // the A operand is the previous kernel's output
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
  } else if (warp == TC_MMA_WARP) {
    // ================================================================== MMA issuer
    // whole warp walks the loop; one elected lane issues / commits; descriptors = base + compile-time offsets
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc_i8(TILE_M, C::N_TILE);
    const uint64_t a_desc0 = make_kmajor_desc<C::KC>(smem_u32(a_smem), 8 * C::KC);
    const uint64_t b_desc0 = make_kmajor_desc<C::KC>(smem_u32(b_smem), 8 * C::KC);
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
        if (leader) {
          const uint64_t da0 = a_desc0 + (uint64_t)((stage * C::A_BYTES) >> 4);
          const uint64_t db0 = b_desc0 + (uint64_t)(((B_RESIDENT ? j : (int)stage) * C::B_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < C::KC / 32; ++k)
            tc_mma_i8(d_tmem, da0 + (uint64_t)((k * 32) >> 4), db0 + (uint64_t)((k * 32) >> 4), idesc,
                      (j | k) != 0 ? 1u : 0u);
          tc_commit(empty_bar + stage);  // smem slot reusable once these MMAs have read it
          if (j == C::NCHUNK - 1) tc_commit(tmem_full_bar + slot);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ================================================================== epilogue warps
    const int quarter = warp & 3;             // TMEM lane quarter this warp may read
    const int half = warp >> 2;               // which half of the N columns
    const int row = quarter * 32 + lane;      // accumulator row == pixel of the tile
    const int et = threadIdx.x;               // 0..255 among epilogue threads
    const bool fast = args.bounded != 0;
    // one 32-column chunk per warp (N = 64): its per-channel constants never change, keep them in registers
    constexpr bool REG_CONSTS = C::CHUNKS_PER_WARP == 1 && C::N_TILES == 1;
    float4 mu_r[REG_CONSTS ? 8 : 1], bd_r[REG_CONSTS ? 8 : 1];
    if constexpr (REG_CONSTS) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        mu_r[g] = *reinterpret_cast<const float4*>(s_mult + half * C::COLS_PER_WARP + 4 * g);
        bd_r[g] = *reinterpret_cast<const float4*>(s_bdiv + half * C::COLS_PER_WARP + 4 * g);
      }
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_tile = tile / C::N_TILES;
      const int n0 = (tile % C::N_TILES) * C::N_TILE + half * C::COLS_PER_WARP;
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
      const int32_t* cm_row = s_cm + cfg * C::CM_STRIDE + n0;

      mbar_wait(tmem_full_bar + slot, acc_phase);
      tc_fence_after();
      const uint32_t t_addr =
          tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * C::N_TILE + half * C::COLS_PER_WARP;
#pragma unroll 1
      for (int ch = 0; ch < C::CHUNKS_PER_WARP; ++ch) {
        const int c0 = ch * 32;
        uint32_t v[32];
        tmem_ld_32x32(t_addr + c0, v);
        tmem_ld_wait();
        if (ch == C::CHUNKS_PER_WARP - 1) {  // accumulator slot drained by this warp: hand it back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar + slot);
        }
        uint32_t packed[8];
        if constexpr (REG_CONSTS) {
          requant_chunk32<CHECK>(v, reinterpret_cast<const int4*>(cm_row + c0), bd_r, mu_r, fast, args.zp_out, args.lo,
                                 packed);
        } else {
          requant_chunk32<CHECK>(v, reinterpret_cast<const int4*>(cm_row + c0),
                                 reinterpret_cast<const float4*>(s_bdiv + n0 + c0),
                                 reinterpret_cast<const float4*>(s_mult + n0 + c0), fast, args.zp_out, args.lo, packed);
        }
        if constexpr (POOL) {
          const int j0 = (half * C::COLS_PER_WARP + c0) >> 4;
          *reinterpret_cast<uint4*>(staging + staging_off<C::N_TILE>(row, j0)) =
              make_uint4(packed[0], packed[1], packed[2], packed[3]);
          *reinterpret_cast<uint4*>(staging + staging_off<C::N_TILE>(row, j0 + 1)) =
              make_uint4(packed[4], packed[5], packed[6], packed[7]);
        } else if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(out + c0);
          dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        }
      }
      if constexpr (POOL) {
        // tile staged -> 2x2 max over pixel windows -> pooled NHWC store; 16 bytes (16 channels) per unit
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
        constexpr int CH16 = C::N_TILE / 16;
        constexpr int PW = IMG / 2, PR = C::ROWS / 2;       // pooled columns / rows per image in this tile
        constexpr int UNITS = C::NB * PR * PW * CH16;       // == 32 * CH16
        const int img_base = (m_tile / C::TILES_PER_IMG) * C::NB;
        const int prow0 = (m_tile % C::TILES_PER_IMG) * PR;
#pragma unroll
        for (int u = et; u < UNITS; u += 32 * TC_EPI_WARPS) {
          const int j = u % CH16;
          int pp = u / CH16;
          const int pw = pp % PW;
          pp /= PW;
          const int pr = pp % PR;
          const int nb = pp / PR;
          const int p00 = (nb * C::ROWS + 2 * pr) * IMG + 2 * pw;
          const uint4 a = *reinterpret_cast<const uint4*>(staging + staging_off<C::N_TILE>(p00, j));
          const uint4 b = *reinterpret_cast<const uint4*>(staging + staging_off<C::N_TILE>(p00 + 1, j));
          const uint4 c = *reinterpret_cast<const uint4*>(staging + staging_off<C::N_TILE>(p00 + IMG, j));
          const uint4 d = *reinterpret_cast<const uint4*>(staging + staging_off<C::N_TILE>(p00 + IMG + 1, j));
          uint4 o;
          o.x = max4_u8x4(a.x, b.x, c.x, d.x);
          o.y = max4_u8x4(a.y, b.y, c.y, d.y);
          o.z = max4_u8x4(a.z, b.z, c.z, d.z);
          o.w = max4_u8x4(a.w, b.w, c.w, d.w);
          const int64_t img = img_base + nb;
          if (img < args.m_rows) {
            uint8_t* dst = args.y + ((img * (IMG / 2) + prow0 + pr) * (int64_t)PW + pw) * COUT +
                           (tile % C::N_TILES) * C::N_TILE + j * 16;
            *reinterpret_cast<uint4*>(dst) = o;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");  // staging free for the next tile
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// --------------------------------------------------------------------------------------------------------- host side
template <int IMG, int CIN, int COUT, bool B_RESIDENT, bool POOL, bool CHECK = true, int NT = 0>
static int launch_tc(const uint8_t* x, uint8_t* y, int64_t m_rows, const int8_t* w, const int32_t* corr,
                     const b200q_requant& rq, cudaStream_t stream) {
  using C = TcCfg<IMG, CIN, COUT, B_RESIDENT, POOL, NT>;
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
  if constexpr (CHECK && COUT == 64) {  // epilogue-critical layers: drop the per-element range test when it is provably idle
    if ((rq.flags & B200Q_RQ_BOUNDED) && (rq.flags & B200Q_RQ_ACC22))
      return launch_tc<IMG, CIN, COUT, B_RESIDENT, POOL, false, NT>(x, y, m_rows, w, corr, rq, stream);
  }
  auto kernel = igemm_tc_kernel<IMG, CIN, COUT, B_RESIDENT, POOL, CHECK, NT>;
  static uint64_t attr_mask = 0;  // per template instantiation
  if (int arc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), C::SMEM_BYTES, &attr_mask)) return arc;
  TcArgs args{y, rq.mult, rq.bdiv, corr, m_rows, num_m_tiles, rq.zp_out, rq.relu ? rq.zp_out : 0,
              (rq.flags & B200Q_RQ_BOUNDED) ? 1 : 0};
  const int num_tiles = num_m_tiles * C::N_TILES;
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  return launch_kernel("igemm_tc_kernel", kernel, grid, TC_THREADS, C::SMEM_BYTES, stream, map_a, map_b, args);
}

}  // namespace b200q

using namespace b200q;

// -DB200Q_DEV builds only: B200Q_TC_STREAM_WEIGHTS=1 forces the streamed-weights variant for layers that default to resident weights
// (bring-up / A-B testing only).
static bool force_streamed() {
#ifndef B200Q_DEV
  return false;
#endif
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200Q_TC_STREAM_WEIGHTS");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// -DB200Q_DEV builds only: B200Q_NO_CTA2=1 keeps conv2 on the single-CTA kernel (A-B timing only).
static bool no_cta2() {
#ifndef B200Q_DEV
  return false;
#endif
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200Q_NO_CTA2");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// -DB200Q_DEV builds only: B200Q_NO_SMALL=1 disables the small-batch kernel selection (A-B timing only).
static bool small_batch_off() {
#ifndef B200Q_DEV
  return false;
#endif
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200Q_NO_SMALL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// Largest batch that takes the CUDA-core kernel of conv_small.cu.  -DB200Q_DEV builds only: B200Q_TINY_MAX_B overrides
// it (0 disables the kernel; A-B timing only).
constexpr int CONV_TINY_MAX_B = 2;
static int tiny_max_b() {
#ifndef B200Q_DEV
  return CONV_TINY_MAX_B;
#endif
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200Q_TINY_MAX_B");
    v = e ? atoi(e) : CONV_TINY_MAX_B;
  }
  return v;
}

// -DB200Q_DEV builds only: B200Q_NO_HALO=1 routes the cin=64 layers through the shifted-TMA kernel as well (A-B testing only).
static bool no_halo() {
#ifndef B200Q_DEV
  return false;
#endif
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200Q_NO_HALO");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

extern "C" int b200q_conv3x3_tc(const uint8_t* x, uint8_t* y, int64_t b, const b200q_conv3x3* L, int pool2x2,
                                void* stream) {
  B200Q_REQUIRE(L && ((x && y) || b == 0), "conv3x3_tc: null pointer");
  B200Q_REQUIRE(L->w && L->corr && L->rq.mult && L->rq.bdiv, "conv3x3_tc: unpacked layer");
  B200Q_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)L->w % 16 == 0,
                "conv3x3_tc: buffers must be 16-byte aligned");
  if (b == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const bool streamed = force_streamed();
  const bool pool = pool2x2 != 0;
  // A handful of images: one round trip per layer on the CUDA cores (conv_small.cu).
  if (b <= tiny_max_b()) {
    int rc = 0;
    if (conv3x3_tiny_dispatch(x, y, b, L, pool, s, &rc) == 0) return rc;
  }
  // Small batches (the latency-bound regime the whole-network executor serves): the band-resident kernels below give a
  // whole image (or four) to ONE CTA, so at batch 1 a layer is a serial walk over its tiles on one SM.  When the layer
  // cut into 128-pixel x 64-channel tiles still fits in about two waves of CTAs, it runs as that many independent CTAs
  // instead (shifted-TMA kernel, N tile 64, streamed weights).  Same arithmetic, bit-identical results.
  {
    const int64_t m_tiles = L->img == 32 ? b * 8 : L->img == 16 ? b * 2 : (b + 1) / 2;
    const bool small = !small_batch_off() && m_tiles * (L->cout / 64) <= 2 * (int64_t)num_sms();
#define B200Q_SMALL_CASE(IMG_, CIN_, COUT_)                                                                       \
  if (small && L->img == IMG_ && L->cin == CIN_ && L->cout == COUT_) {                                            \
    if (pool) return launch_tc<IMG_, CIN_, COUT_, false, true, true, 64>(x, y, b, L->w, L->corr, L->rq, s);       \
    return launch_tc<IMG_, CIN_, COUT_, false, false, true, 64>(x, y, b, L->w, L->corr, L->rq, s);                \
  }
    B200Q_SMALL_CASE(32, 64, 64)
    B200Q_SMALL_CASE(16, 64, 128)
    B200Q_SMALL_CASE(16, 128, 128)
    B200Q_SMALL_CASE(8, 128, 256)
    B200Q_SMALL_CASE(8, 256, 256)
#undef B200Q_SMALL_CASE
  }
  if (!no_halo()) {
    int rc = 0;
    if (!no_cta2() && conv3x3_halo2_dispatch(x, y, b, L, pool, s, &rc) == 0) return rc;
    if (conv3x3_halo_dispatch(x, y, b, L, pool, s, &rc) == 0) return rc;
    if (conv3x3_pair_dispatch(x, y, b, L, pool, s, &rc) == 0) return rc;
  }
#define B200Q_TC_CASE(IMG_, CIN_, COUT_, RES_)                                                                  \
  if (L->img == IMG_ && L->cin == CIN_ && L->cout == COUT_) {                                                   \
    if (RES_ && !streamed) {                                                                                    \
      if (pool) return launch_tc<IMG_, CIN_, COUT_, RES_, true>(x, y, b, L->w, L->corr, L->rq, s);              \
      return launch_tc<IMG_, CIN_, COUT_, RES_, false>(x, y, b, L->w, L->corr, L->rq, s);                       \
    }                                                                                                           \
    if (pool) return launch_tc<IMG_, CIN_, COUT_, false, true>(x, y, b, L->w, L->corr, L->rq, s);               \
    return launch_tc<IMG_, CIN_, COUT_, false, false>(x, y, b, L->w, L->corr, L->rq, s);                        \
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
  if (L->k == 4096 && L->n == 512) {
    // few rows: eight 64-column CTA tiles per 128 rows instead of two 256-column ones (each CTA then streams 256 KB of
    // the 2 MB weight matrix instead of 1 MB)
    if (!small_batch_off() && (b + TILE_M - 1) / TILE_M * 8 <= (int64_t)num_sms())
      return launch_tc<0, 4096, 512, false, false, true, 64>(x, y, b, L->w, L->corr, L->rq, s);
    return launch_tc<0, 4096, 512, false, false>(x, y, b, L->w, L->corr, L->rq, s);
  }
  if (L->k == 512 && L->n == 64) return launch_tc<0, 512, 64, false, false>(x, y, b, L->w, L->corr, L->rq, s);
  set_error("linear_tc: unsupported geometry k=%d n=%d", L->k, L->n);
  return B200Q_ERR_INVALID_ARG;
}
