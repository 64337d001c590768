// Halo-resident tcgen05 implicit-GEMM 3x3 convolution for the 64-input-channel layers (conv2, conv3).
//
// The shifted-TMA formulation (igemm_tc.cu) re-reads every activation tile from L2 nine times (once per filter tap);
// on B200 the L2->SM path sustains about what HBM does, so those layers were L2-bound at ~4x their MMA time.  Here
// a band of NBI whole images is fetched ONCE per CTA iteration by a single 4-D TMA box that starts at column -1 and
// row -1 of each image: out-of-image elements are zero-filled, which lays the band out in shared memory as one long
// sequence of pixels with pitch P = IMG+1 (one zero column shared by neighbouring rows, one zero row shared by
// neighbouring images).  In that sequence the input of filter tap (kh,kw) for output position q is simply position
// q + (kh-1)*P + (kw-1), so the nine A operands of a 128-position tile are the SAME shared-memory array viewed through
// UMMA descriptors whose start address is shifted by a whole number of 64-byte rows (the swizzle is a function of the
// absolute shared-memory address, so row-shifted descriptors read TMA-written data correctly; tools/probe_shift.cu).
// Tiles are 128 consecutive positions; the pad positions inside a tile produce rows that are simply not stored
// (3-11 % of the MMA work).  Weights stay resident in shared memory; accumulators are double-buffered in TMEM.
//
// Warp roles (320 threads): warps 0..7 = epilogue (two per TMEM lane quarter, half of the channels each), warp 8 = TMA
// producer, warp 9 = MMA issuer / TMEM owner.  The issuer has the highest warp id on purpose: the sub-partition
// arbiter prefers higher warp ids, and a starved issuer stalls the tensor pipe for everyone.  Epilogue arithmetic: common.cuh requant_chunk32 (fbgemm-exact).
// With POOL the requantised tile goes to a two-tile ring in shared memory and every 2x2 window whose last pixel lies
// in the current tile is max-reduced and stored (aten::quantized_max_pool2d fused; max commutes with the monotone
// requantisation).
#include "common.cuh"

namespace b200q {

constexpr int HALO_EPI_WARPS = 8;
constexpr int HALO_THREADS = 64 + 32 * HALO_EPI_WARPS;
constexpr int HALO_TMA_WARP = HALO_EPI_WARPS, HALO_MMA_WARP = HALO_EPI_WARPS + 1;

template <int IMG, int COUT, int NBI, bool POOL>
struct HaloCfg {
  static constexpr int CIN = 64;                 // bytes per pixel row == swizzle span (SWIZZLE_64B)
  static constexpr int P = IMG + 1;              // pitch of the padded pixel sequence
  static constexpr int POS_PER_IMG = (IMG + 1) * P;
  static constexpr int BOX_POS = NBI * POS_PER_IMG;
  static constexpr int Q0 = P + 1;               // position of pixel (0,0) of the band's first image
  static constexpr int Q_LAST = (NBI - 1) * POS_PER_IMG + IMG * P + IMG;
  static constexpr int TILES = (Q_LAST - Q0 + 1 + 127) / 128;
  static constexpr int A_POS = Q0 + 128 * TILES + P + 1;           // positions any tap of any tile may touch
  static constexpr int BOX_BYTES = BOX_POS * CIN;
  static constexpr int A_BYTES = (A_POS * CIN + 1023) / 1024 * 1024;
  static constexpr int W_TAP_BYTES = COUT * CIN;
  static constexpr int W_BYTES = 9 * W_TAP_BYTES;
  static constexpr int CM_STRIDE = COUT + 4;
  static constexpr int TABLE_BYTES = (9 * CM_STRIDE + 2 * COUT) * 4;
  static constexpr int STAGING_BYTES = POOL ? 2 * 128 * COUT : 0;
  // look-up tables (identical for every band): per tile row -> {valid, border class, pixel index in the band};
  // per tile -> list of 2x2 windows completed by that tile {ring row of the bottom-right pixel, pooled pixel index}
  static constexpr int MAX_CORNERS = 39;  // +1 count word = 40 words per tile (keeps the mbarriers behind it 8-byte aligned)
  static constexpr int LUT_BYTES = TILES * 128 * 4 + (POOL ? TILES * (MAX_CORNERS + 1) * 4 : 0);
  static constexpr int SMEM_BYTES =
      2 * A_BYTES + W_BYTES + STAGING_BYTES + TABLE_BYTES + LUT_BYTES + 256 /*barriers*/ + 1024;
  static constexpr int TMEM_COLS = 2 * COUT;
  static constexpr int COLS_PER_WARP = COUT / 2;
  static constexpr int CHUNKS_PER_WARP = COLS_PER_WARP / 32;
  static_assert(COUT == 64 || COUT == 128, "COUT");
  static_assert(A_POS >= BOX_POS + P + 1, "the zero row below the last image must lie inside the A buffer");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
  static_assert(!POOL || IMG % 2 == 0, "pool");
};

struct HaloArgs {
  uint8_t* y;
  const float* mult;
  const float* bdiv;
  const int32_t* corr;
  int64_t n_img;
  int num_bands;
  int zp_out, lo;
  int bounded;
};

template <int N>
__device__ __forceinline__ int halo_staging_off(int row, int j) {  // row in [0,256): two-tile ring
  const int f = (N == 64) ? ((row >> 1) & 3) : (row & 7);
  return row * N + ((j ^ f) << 4);
}

template <int IMG, int COUT, int NBI, bool POOL, bool CHECK>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                 const HaloArgs args) {
  using C = HaloCfg<IMG, COUT, NBI, POOL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                                  // 2 x A_BYTES
  uint8_t* w_smem = a_smem + 2 * C::A_BYTES;               // 9 x [COUT][64]
  uint8_t* staging = w_smem + C::W_BYTES;                  // POOL: [2][128][COUT]
  int32_t* s_cm = reinterpret_cast<int32_t*>(staging + C::STAGING_BYTES);
  float* s_mult = reinterpret_cast<float*>(s_cm + 9 * C::CM_STRIDE);
  float* s_bdiv = s_mult + COUT;
  uint32_t* s_rowlut = reinterpret_cast<uint32_t*>(s_bdiv + COUT);     // [TILES][128]
  uint32_t* s_corner = s_rowlut + C::TILES * 128;                       // POOL: [TILES][1 + MAX_CORNERS]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_corner + (POOL ? C::TILES * (C::MAX_CORNERS + 1) : 0));  // [2]
  uint64_t* empty_bar = full_bar + 2;                                  // [2] band consumed by the MMAs
  uint64_t* w_bar = empty_bar + 2;                                     // weights landed
  uint64_t* tmem_full_bar = w_bar + 1;                                 // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                        // [2]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == HALO_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
      mbar_init(tmem_full_bar + i, 1);
      mbar_init(tmem_empty_bar + i, HALO_EPI_WARPS);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == HALO_MMA_WARP) {
    tmem_alloc(tmem_base_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp < HALO_EPI_WARPS) {
    const int t = threadIdx.x;
    for (int i = t; i < 9 * COUT; i += 32 * HALO_EPI_WARPS)
      s_cm[(i / COUT) * C::CM_STRIDE + i % COUT] = (int32_t)(MAGIC_BITS - (uint32_t)__ldg(args.corr + i));
    for (int i = t; i < COUT; i += 32 * HALO_EPI_WARPS) {
      s_mult[i] = __ldg(args.mult + i);
      s_bdiv[i] = __ldg(args.bdiv + i);
    }
    for (int i = t; i < C::TILES * 128; i += 32 * HALO_EPI_WARPS) {
      const int q = C::Q0 + i;  // tiles are consecutive: position of row (i % 128) of tile (i / 128)
      const int bi = q / C::POS_PER_IMG;
      const int rem = q - bi * C::POS_PER_IMG;
      const int h = rem / C::P - 1;
      const int w = rem - (h + 1) * C::P - 1;
      const bool ok = bi < NBI && h >= 0 && w >= 0;
      const int cfg = (h <= 0 ? 0 : (h == IMG - 1 ? 2 : 1)) * 3 + (w <= 0 ? 0 : (w == IMG - 1 ? 2 : 1));
      s_rowlut[i] = ok ? (0x80000000u | ((uint32_t)cfg << 16) | (uint32_t)((bi * IMG + h) * IMG + w)) : 0u;
    }
    if constexpr (POOL) {
      if (t < C::TILES) {  // one thread per tile compacts the windows whose bottom-right pixel lies in that tile
        uint32_t* list = s_corner + t * (C::MAX_CORNERS + 1);
        int n = 0;
        for (int r = 0; r < 128; ++r) {
          const int q = C::Q0 + 128 * t + r;
          const int bi = q / C::POS_PER_IMG;
          const int rem = q - bi * C::POS_PER_IMG;
          const int h = rem / C::P - 1;
          const int w = rem - (h + 1) * C::P - 1;
          if (bi < NBI && h > 0 && w > 0 && (h & 1) && (w & 1) && n < C::MAX_CORNERS)
            list[1 + n++] = ((uint32_t)((t & 1) * 128 + r) << 16) |
                            (uint32_t)((bi * (IMG / 2) + (h >> 1)) * (IMG / 2) + (w >> 1));
        }
        list[0] = (uint32_t)n;
      }
    }
    // everything behind the TMA box (zero row below the last image + slack read only by discarded rows) stays zero
    for (int buf = 0; buf < 2; ++buf) {
      uint4* tail = reinterpret_cast<uint4*>(a_smem + buf * C::A_BYTES + C::BOX_BYTES);
      for (int i = t; i < (C::A_BYTES - C::BOX_BYTES) / 16; i += 32 * HALO_EPI_WARPS) tail[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();  // generic-proxy zeros -> visible to the tensor core's async-proxy reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == HALO_TMA_WARP) {
    // ================================================================== TMA producer
    if (lane == 0) {
      mbar_expect_tx(w_bar, C::W_BYTES);
      for (int tap = 0; tap < 9; ++tap)
        tma_load_2d(w_smem + tap * C::W_TAP_BYTES, &map_w, w_bar, tap * C::CIN, 0);
      int it = 0;
      for (int band = blockIdx.x; band < args.num_bands; band += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(empty_bar + buf, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(full_bar + buf, C::BOX_BYTES);
        tma_load_4d(a_smem + buf * C::A_BYTES, &map_a, full_bar + buf, 0, -1, -1, band * NBI);
      }
    }
  } else if (warp == HALO_MMA_WARP) {
    // ================================================================== MMA issuer
    // The whole warp walks the loop (uniform control flow, waits included); one elected lane issues the MMAs and
    // commits.  Descriptors are built once per band / tile; per MMA only compile-time offsets are added, so the
    // issue rate stays far above the 32..64 cycles an MMA takes.
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc_i8(128, COUT);
    mbar_wait(w_bar, 0);
    const uint64_t w_desc0 = make_kmajor_desc<C::CIN>(smem_u32(w_smem), 8 * C::CIN);
    int it = 0, acc_it = 0;
    for (int band = blockIdx.x; band < args.num_bands; band += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(full_bar + buf, (it >> 1) & 1);
      tc_fence_after();
      const uint64_t a_desc0 = make_kmajor_desc<C::CIN>(smem_u32(a_smem + buf * C::A_BYTES), 8 * C::CIN);
      for (int t = 0; t < C::TILES; ++t, ++acc_it) {
        const uint32_t slot = acc_it & 1;
        mbar_wait(tmem_empty_bar + slot, ((acc_it >> 1) & 1) ^ 1);
        tc_fence_after();
        if (leader) {
          const uint32_t d_tmem = tmem_base + slot * COUT;
          const uint64_t a_tile = a_desc0 + (uint64_t)(((C::Q0 + 128 * t - C::P - 1) * C::CIN) >> 4);  // tap (0,0)
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int k = 0; k < C::CIN / 32; ++k) {
              const uint64_t da = a_tile + (uint64_t)((((tap / 3) * C::P + (tap % 3)) * C::CIN + k * 32) >> 4);
              const uint64_t db = w_desc0 + (uint64_t)((tap * C::W_TAP_BYTES + k * 32) >> 4);
              tc_mma_i8(d_tmem, da, db, idesc, (tap | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(tmem_full_bar + slot);
        }
        __syncwarp();
      }
      if (leader) tc_commit(empty_bar + buf);  // arrives when every MMA that reads this band has completed
      __syncwarp();
    }
  } else {
    // ================================================================== epilogue warps
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int row = quarter * 32 + lane;
    const int et = threadIdx.x;
    const int n0 = half * C::COLS_PER_WARP;
    const bool fast = args.bounded != 0;
    constexpr bool REG_CONSTS = C::CHUNKS_PER_WARP == 1;
    float4 mu_r[REG_CONSTS ? 8 : 1], bd_r[REG_CONSTS ? 8 : 1];
    if constexpr (REG_CONSTS) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        mu_r[g] = *reinterpret_cast<const float4*>(s_mult + n0 + 4 * g);
        bd_r[g] = *reinterpret_cast<const float4*>(s_bdiv + n0 + 4 * g);
      }
    }
    int acc_it = 0;
    for (int band = blockIdx.x; band < args.num_bands; band += gridDim.x) {
      const int64_t img0 = (int64_t)band * NBI;
      for (int t = 0; t < C::TILES; ++t, ++acc_it) {
        const uint32_t slot = acc_it & 1;
        // row of the tile -> {valid, border class, pixel index in the band}; pad positions / images past the batch are
        // computed but not stored
        const uint32_t e = s_rowlut[t * 128 + row];
        const int pix = (int)(e & 0xffffu);
        const bool valid = (e >> 31) && img0 + pix / (IMG * IMG) < args.n_img;
        const int32_t* cm_row = s_cm + ((e >> 16) & 0xf) * C::CM_STRIDE + n0;

        mbar_wait(tmem_full_bar + slot, (acc_it >> 1) & 1);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * COUT + n0;
#pragma unroll 1
        for (int ch = 0; ch < C::CHUNKS_PER_WARP; ++ch) {
          const int c0 = ch * 32;
          uint32_t v[32];
          tmem_ld_32x32(t_addr + c0, v);
          tmem_ld_wait();
          if (ch == C::CHUNKS_PER_WARP - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar + slot);
          }
          uint32_t packed[8];
          if constexpr (REG_CONSTS) {
            requant_chunk32<CHECK>(v, reinterpret_cast<const int4*>(cm_row + c0), bd_r, mu_r, fast, args.zp_out,
                                   args.lo, packed);
          } else {
            requant_chunk32<CHECK>(v, reinterpret_cast<const int4*>(cm_row + c0),
                                   reinterpret_cast<const float4*>(s_bdiv + n0 + c0),
                                   reinterpret_cast<const float4*>(s_mult + n0 + c0), fast, args.zp_out, args.lo, packed);
          }
          if constexpr (POOL) {
            const int srow = (t & 1) * 128 + row;
            const int j0 = (n0 + c0) >> 4;
            *reinterpret_cast<uint4*>(staging + halo_staging_off<COUT>(srow, j0)) =
                make_uint4(packed[0], packed[1], packed[2], packed[3]);
            *reinterpret_cast<uint4*>(staging + halo_staging_off<COUT>(srow, j0 + 1)) =
                make_uint4(packed[4], packed[5], packed[6], packed[7]);
          } else if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(args.y + (img0 * (IMG * IMG) + pix) * (int64_t)COUT + n0 + c0);
            dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          }
        }
        if constexpr (POOL) {
          // windows whose bottom-right pixel (odd row, odd column) lies in this tile are complete: the other three
          // pixels are at most P+1 positions back, i.e. in this tile or the previous one (ring of two tiles)
          asm volatile("bar.sync 1, %0;" ::"n"(32 * HALO_EPI_WARPS) : "memory");
          constexpr int CH16 = COUT / 16;
          const uint32_t* list = s_corner + t * (C::MAX_CORNERS + 1);
          const int units = (int)list[0] * CH16;
          for (int u = et; u < units; u += 32 * HALO_EPI_WARPS) {
            const int j = u % CH16;
            const uint32_t ce = list[1 + u / CH16];
            const int s11 = (int)(ce >> 16);   // ring row of the bottom-right pixel; neighbours wrap modulo 256
            const int ppix = (int)(ce & 0xffffu);
            if (img0 + ppix / ((IMG / 2) * (IMG / 2)) < args.n_img) {
              const uint4 a = *reinterpret_cast<const uint4*>(staging + halo_staging_off<COUT>((s11 - C::P - 1) & 255, j));
              const uint4 b = *reinterpret_cast<const uint4*>(staging + halo_staging_off<COUT>((s11 - C::P) & 255, j));
              const uint4 c = *reinterpret_cast<const uint4*>(staging + halo_staging_off<COUT>((s11 - 1) & 255, j));
              const uint4 d = *reinterpret_cast<const uint4*>(staging + halo_staging_off<COUT>(s11, j));
              uint4 o;
              o.x = max4_u8x4(a.x, b.x, c.x, d.x);
              o.y = max4_u8x4(a.y, b.y, c.y, d.y);
              o.z = max4_u8x4(a.z, b.z, c.z, d.z);
              o.w = max4_u8x4(a.w, b.w, c.w, d.w);
              uint8_t* dst = args.y + (img0 * ((IMG / 2) * (IMG / 2)) + ppix) * (int64_t)COUT + j * 16;
              *reinterpret_cast<uint4*>(dst) = o;
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * HALO_EPI_WARPS) : "memory");
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == HALO_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int IMG, int COUT, int NBI, bool POOL, bool CHECK = true>
static int launch_halo(const uint8_t* x, uint8_t* y, int64_t n_img, const int8_t* w, const int32_t* corr,
                       const b200q_requant& rq, cudaStream_t stream) {
  using C = HaloCfg<IMG, COUT, NBI, POOL>;
  if constexpr (CHECK && COUT == 64) {
    if ((rq.flags & B200Q_RQ_BOUNDED) && (rq.flags & B200Q_RQ_ACC22))
      return launch_halo<IMG, COUT, NBI, POOL, false>(x, y, n_img, w, corr, rq, stream);
  }
  CUtensorMap map_a, map_w;
  {
    const uint64_t dims[4] = {(uint64_t)C::CIN, (uint64_t)IMG, (uint64_t)IMG, (uint64_t)n_img};
    const uint64_t strides[3] = {(uint64_t)C::CIN, (uint64_t)IMG * C::CIN, (uint64_t)IMG * IMG * C::CIN};
    const uint32_t box[4] = {(uint32_t)C::CIN, (uint32_t)C::P, (uint32_t)(IMG + 1), (uint32_t)NBI};
    int rc = encode_tensor_map(&map_a, x, 4, dims, strides, box, C::CIN);
    if (rc) return rc;
  }
  {
    const uint64_t ktot = 9ull * C::CIN;
    const uint64_t dims[2] = {ktot, (uint64_t)COUT};
    const uint64_t strides[1] = {ktot};
    const uint32_t box[2] = {(uint32_t)C::CIN, (uint32_t)COUT};
    int rc = encode_tensor_map(&map_w, w, 2, dims, strides, box, C::CIN);
    if (rc) return rc;
  }
  auto kernel = conv_halo_kernel<IMG, COUT, NBI, POOL, CHECK>;
  static bool attr_set = false;
  if (!attr_set) {
    B200Q_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int num_bands = (int)((n_img + NBI - 1) / NBI);
  HaloArgs args{y, rq.mult, rq.bdiv, corr, n_img, num_bands, rq.zp_out, rq.relu ? rq.zp_out : 0,
                (rq.flags & B200Q_RQ_BOUNDED) ? 1 : 0};
  const int grid = num_bands < num_sms() ? num_bands : num_sms();
  kernel<<<grid, HALO_THREADS, C::SMEM_BYTES, stream>>>(map_a, map_w, args);
  return launched("conv_halo_kernel");
}

// Entry used by b200q_conv3x3_tc for the geometries this kernel covers; returns 1 when the geometry is not handled.
int conv3x3_halo_dispatch(const uint8_t* x, uint8_t* y, int64_t b, const b200q_conv3x3* L, bool pool, cudaStream_t s,
                          int* rc) {
  if (L->cin != 64) return 1;
  if (L->img == 32 && L->cout == 64) {
    *rc = pool ? launch_halo<32, 64, 1, true>(x, y, b, L->w, L->corr, L->rq, s)
               : launch_halo<32, 64, 1, false>(x, y, b, L->w, L->corr, L->rq, s);
    return 0;
  }
  if (L->img == 16 && L->cout == 128 && !pool) {
    *rc = launch_halo<16, 128, 3, false>(x, y, b, L->w, L->corr, L->rq, s);
    return 0;
  }
  return 1;
}

}  // namespace b200q
