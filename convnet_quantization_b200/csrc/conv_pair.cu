// Pair-interleaved halo kernel for the 8x8 layers (conv5: cin=128, conv6: cin=256; cout=256).
//
// The shifted-TMA kernel (igemm_tc.cu) that served these layers first re-read every activation tile from L2 nine times
// and re-streamed the weights (295 / 590 KB) for every 128-pixel tile: ~96 B/clk/SM of L2 traffic, about 1.5x what the
// L2 -> SM path sustains, so both layers ran at 55-65 % of their tensor-pipe time.  This kernel keeps conv_halo.cu's idea
// (activations fetched once into a padded pixel sequence in shared memory; the nine taps are row-shifted UMMA
// descriptors over the same array) and adds what the small images and the large weight matrices need:
//
// * An MMA tile (M = 128) is a PAIR of 8x8 images whose rows are interleaved in the sequence: position
//   (h+1)*18 + 9*image + (w+1), i.e. [pad, image 0 row h, pad, image 1 row h] per sequence row of pitch 18.  The 8-pixel
//   row groups of the pair then follow each other at a constant stride of 9 positions (the descriptor's SBO), and a
//   filter tap is a shift of kh*18 + kw positions.  Accumulator row 16*h + 8*image + w.
// * A CTA computes ONE HALF of the output channels (N = 128: the tensor pipe's full rate; blockIdx & 1 selects the half)
//   for a band of two pairs (4 images) at a time, so every weight chunk fetched from L2 feeds two MMAs' worth of tiles:
//   147 / 295 KB of weights per band of 2 x 36 / 72 MMAs = 32 B/clk/SM.  The two channel halves of a band are computed
//   by different CTAs, which costs a second read of the (small) activations and keeps the epilogue constants of a
//   thread fixed for the lifetime of the CTA.
// * Shared-memory rows are KC = 128 bytes (the widest swizzle span; weight chunks of 64-byte rows made the TMA the
//   bottleneck of conv5: 0.28 ms instead of 0.14).  conv6's 256 input channels are therefore two K halves held in
//   separate arrays.  Its MMAs run K half 0 first, then K half 1, so with ONE buffer per half (two would not fit) the
//   loader refills half 0 with the next band while half 1 is being consumed, and vice versa.  conv5 has one K "half"
//   and double-buffers it.
// * Weights stream through a ring of [128 channels][KC] chunks (one per K half and tap) filled by 5-D TMA boxes that
//   deliver the rows in the epilogue's channel permutation (epilogue16.cuh).
// * Warps: 8 epilogue (epilogue16.cuh, 16 channels per thread), activation loader, weight producer, two MMA issuers
//   (one per tile of the band; see conv_halo.cu for why two).
#include "common.cuh"
#include "epilogue16.cuh"

namespace b200q {

constexpr int PAIR_IMG = 8, PAIR_COUT = 256, PAIR_N = 128;  // N per CTA
constexpr int PAIR_EPI_WARPS = 8, PAIR_NCH = 16;
constexpr int PAIR_LOAD_WARP = PAIR_EPI_WARPS, PAIR_W_WARP = PAIR_EPI_WARPS + 1, PAIR_MMA_WARP = PAIR_EPI_WARPS + 2;
constexpr int PAIR_THREADS = 32 * (PAIR_EPI_WARPS + 4);
constexpr int PAIR_SLOTS = 4;   // TMEM accumulator slots of 128 columns
constexpr int PAIR_T = 2;       // pairs (= tiles) per band

template <int CIN>
struct PairCfg {
  static constexpr int KC = 128;                      // bytes per row of one K half == swizzle span
  static constexpr int KH = CIN / KC;                 // K halves: 1 (conv5) or 2 (conv6)
  static constexpr int P = 2 * PAIR_IMG + 2;          // 18: sequence pitch of an interleaved row
  static constexpr int PAIR_POS = (PAIR_IMG + 2) * P; // 180: pad row, 8 rows, pad row
  static constexpr int A_POS = PAIR_T * PAIR_POS + 8; // + the positions the last taps of the last tile reach
  static constexpr int A_BYTES = (A_POS * KC + 1023) / 1024 * 1024;   // one K half
  static constexpr int ABUF = KH == 1 ? 2 : 1;                         // band buffers per K half
  static constexpr int B_BYTES = PAIR_N * KC;                          // one weight chunk: (K half, tap)
  static constexpr int STAGES = 6;
  static constexpr int CHUNKS = KH * 9;                                // weight chunks per band
  static constexpr int SMEM_BYTES = KH * ABUF * A_BYTES + STAGES * B_BYTES + 512 /*barriers*/ + 1024 /*alignment slack*/;
  static constexpr int MMAS_PER_CHUNK = KC / 32;
  static_assert(CIN == 128 || CIN == 256, "CIN");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

struct PairArgs {
  const uint8_t* x;
  uint8_t* y;
  int64_t n_img;
  int num_bands;
  int zp_x, zp_out, lo, bounded;
  int debug;  // B200Q_PAIR_DEBUG bits, honoured only by -DB200Q_DEV builds (timing experiments; results are wrong when
              // set): 2 = no MMAs, 4 = no activation copies, 8 = epilogue only drains, 32 = no weight TMA
};
#ifdef B200Q_DEV
#define PAIR_DBG(bit) ((args.debug & (bit)) != 0)
#else
#define PAIR_DBG(bit) false  // compiled out of the product library
#endif

struct alignas(16) PairConsts {  // the CTA's 128 output channels are [128*half, 128*half + 128)
  int32_t cm[PAIR_COUT];
  float k1[PAIR_COUT];
  float bdiv[PAIR_COUT];
  float mult[PAIR_COUT];
};

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

template <int CIN, bool POOL, bool CHECK>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ PairConsts consts,
                 const PairArgs args) {
  using C = PairCfg<CIN>;
  constexpr int IMG = PAIR_IMG, COUT = PAIR_COUT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int NA = C::KH * C::ABUF;                       // activation arrays: index ab = KH*(band buffer) + K half
  uint8_t* a_smem = smem;                                   // [NA][A_BYTES]
  uint8_t* b_smem = a_smem + NA * C::A_BYTES;               // [STAGES][128][KC]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(b_smem + C::STAGES * C::B_BYTES);  // [NA]
  uint64_t* a_empty = a_full + NA;                          // [NA]
  uint64_t* b_full = a_empty + NA;                          // [STAGES]
  uint64_t* b_empty = b_full + C::STAGES;                   // [STAGES]
  uint64_t* tmem_full_bar = b_empty + C::STAGES;            // [PAIR_SLOTS]
  uint64_t* tmem_empty_bar = tmem_full_bar + PAIR_SLOTS;    // [PAIR_SLOTS]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + PAIR_SLOTS);
  uint32_t* magic_smem = tmem_base_smem + 1;  // holds MAGIC_BITS (epilogue16.cuh epi_init)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nhalf = blockIdx.x & 1;                         // which 128 output channels
  const int band0 = blockIdx.x >> 1, band_step = gridDim.x >> 1;
  pdl_launch_dependents();

  if (warp == PAIR_W_WARP && lane == 0) {
    tma_prefetch_desc(&map_w);
    *magic_smem = MAGIC_BITS;
    for (int i = 0; i < NA; ++i) {
      mbar_init(a_full + i, 32);   // one cp.async-completion arrive per loader lane
      mbar_init(a_empty + i, 2);   // one commit per issuer
    }
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(b_full + i, 1);
      mbar_init(b_empty + i, 2);
    }
    for (int i = 0; i < PAIR_SLOTS; ++i) {
      mbar_init(tmem_full_bar + i, 1);
      mbar_init(tmem_empty_bar + i, PAIR_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == PAIR_MMA_WARP) {
    tmem_alloc(tmem_base_smem, PAIR_SLOTS * PAIR_N);
    tmem_relinquish();
  }
  if (warp < PAIR_EPI_WARPS) {
    // Pad positions hold the activation zero-point for the lifetime of the CTA (the loader never writes them): sequence
    // rows 0 and 9 of every pair, positions 0 and 9 of every row, and everything behind the last pair.
    const uint32_t zp4 = (uint32_t)args.zp_x * 0x01010101u;
    const uint4 zpv = make_uint4(zp4, zp4, zp4, zp4);
    constexpr int CPR = C::KC / 16;
    for (int i = threadIdx.x; i < NA * C::A_POS * CPR; i += 32 * PAIR_EPI_WARPS) {
      const int kh = i / (C::A_POS * CPR), rem = i % (C::A_POS * CPR);  // kh: array index ab
      const int pos = rem / CPR, part = rem % CPR;
      const int in_pair = pos % C::PAIR_POS, rr = in_pair / C::P, cc = in_pair % C::P;
      const bool pad = pos >= PAIR_T * C::PAIR_POS || rr == 0 || rr == IMG + 1 || cc == 0 || cc == IMG + 1;
      // whole rows of one byte value are swizzle-invariant
      if (pad) *reinterpret_cast<uint4*>(a_smem + kh * C::A_BYTES + pos * C::KC + part * 16) = zpv;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  if (warp < 8) {  // pre-bias every accumulator slot (see requant4_prebiased)
    const uint32_t base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    for (int slot = 0; slot < PAIR_SLOTS; ++slot)
      for (int c = 0; c < 64; c += 8) tmem_st_fill8(base + slot * PAIR_N + c, MAGIC_BITS);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == PAIR_LOAD_WARP) {
    // ================================================================== activation loader
    // One warp iteration copies ROWS_PER_IT image rows: lane -> (row-in-iteration, pixel w, 16-byte part).  Destination:
    // position pair*180 + (h+1)*18 + 9*image_in_pair + (w+1), chunk slot part ^ swz(position) (hardware swizzle on
    // absolute addresses; the arrays are 1 KiB aligned).  Everything but the row offset is loop-invariant.
    constexpr int CPR = C::KC / 16;                 // 16-byte chunks per pixel row of a K half
    constexpr int ROWS_PER_IT = 32 / (IMG * CPR) > 0 ? 32 / (IMG * CPR) : 1;
    constexpr int ITS_PER_ROW = IMG * CPR / 32 > 0 ? IMG * CPR / 32 : 1;   // KC = 128: two iterations per image row
    const int part = lane % CPR;
    const int w_it = (lane / CPR) % IMG, w_step = 32 / CPR;                // KC = 128: w = lane/8 + 4*iteration
    const int h_it = lane / (IMG * CPR);                                   // KC = 64 only: 0 (32 lanes = one row)
    pdl_wait();  // the activations are the previous kernel's output
    int it = 0;
    for (int band = band0; band < args.num_bands; band += band_step, ++it) {
      for (int kh = 0; kh < C::KH; ++kh) {
        const int ab = (it % C::ABUF) * C::KH + kh;
        mbar_wait(a_empty + ab, ((it / C::ABUF) & 1) ^ 1);
        const uint32_t a_buf = smem_u32(a_smem + ab * C::A_BYTES);
        for (int bi = 0; bi < 2 * PAIR_T; ++bi) {
          const int64_t img = (int64_t)band * (2 * PAIR_T) + bi;
          if (img >= args.n_img || PAIR_DBG(4)) break;  // stale data: those pixels are never stored
          const uint8_t* src = args.x + img * (int64_t)(IMG * IMG * CIN) + kh * C::KC + part * 16;
          const int pos0 = (bi >> 1) * C::PAIR_POS + (bi & 1) * (IMG + 1) + C::P + 1;  // pixel (0, 0)
#pragma unroll
          for (int i = 0; i < IMG * ITS_PER_ROW / ROWS_PER_IT; ++i) {
            const int h = (i / ITS_PER_ROW) * ROWS_PER_IT + h_it;
            const int w = w_it + (i % ITS_PER_ROW) * w_step;
            const int pos = pos0 + h * C::P + w;
            const int swz = (C::KC == 64) ? ((pos >> 1) & 3) : (pos & 7);
            const uint32_t dst = a_buf + pos * C::KC + ((part ^ swz) << 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (h * IMG + w) * CIN) : "memory");
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(a_full + ab)) : "memory");
      }
    }
  } else if (warp == PAIR_W_WARP) {
    // ================================================================== weight producer (one thread, TMA)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int band = band0; band < args.num_bands; band += band_step) {
        for (int j = 0; j < C::CHUNKS; ++j) {
          const int kh = j / 9, tap = j % 9;
          mbar_wait(b_empty + stage, phase ^ 1);
          if (PAIR_DBG(32)) {
            mbar_arrive(b_full + stage);
            if (++stage == C::STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_expect_tx(b_full + stage, C::B_BYTES);
          uint8_t* dst = b_smem + stage * C::B_BYTES;
          // dims (k, e, Q, b, u): channel = 16*Q + 4*u + 2*b + e; a box is one 64-channel part in epilogue row order
          const int k0 = tap * CIN + kh * C::KC;
          tma_load_5d(dst, &map_w, b_full + stage, k0, 0, 4 * (2 * nhalf), 0, 0);
          tma_load_5d(dst + 64 * C::KC, &map_w, b_full + stage, k0, 0, 4 * (2 * nhalf + 1), 0, 0);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= PAIR_MMA_WARP) {
    // ================================================================== MMA issuers: issuer i owns tile (pair) i
    const int issuer = warp - PAIR_MMA_WARP;
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc_i8(128, PAIR_N);
    const uint64_t b_desc0 = make_kmajor_desc<C::KC>(smem_u32(b_smem), 8 * C::KC);
    // 8-row core groups = the 8 pixels of one image row; consecutive groups are 9 positions apart
    const uint64_t a_desc_tile =
        make_kmajor_desc<C::KC>(smem_u32(a_smem) + issuer * C::PAIR_POS * C::KC, (IMG + 1) * C::KC);  // array 0
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int band = band0; band < args.num_bands; band += band_step, ++it) {
      const int acc_it = it * PAIR_T + issuer;
      const uint32_t slot = acc_it % PAIR_SLOTS;
      mbar_wait(tmem_empty_bar + slot, ((acc_it / PAIR_SLOTS) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + slot * PAIR_N;
      for (int kh = 0; kh < C::KH; ++kh) {
        const int ab = (it % C::ABUF) * C::KH + kh;
        mbar_wait(a_full + ab, (it / C::ABUF) & 1);
        fence_proxy_async_smem();  // cp.async wrote through the generic proxy
        tc_fence_after();
        const uint64_t a_desc_kh = a_desc_tile + (uint64_t)((ab * C::A_BYTES) >> 4);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(b_full + stage, phase);
          tc_fence_after();
          if (leader) {
            const uint64_t da0 = a_desc_kh + (uint64_t)((((tap / 3) * C::P + (tap % 3)) * C::KC) >> 4);
            const uint64_t db0 = b_desc0 + (uint64_t)((stage * C::B_BYTES) >> 4);
#pragma unroll
            for (int k = 0; k < C::MMAS_PER_CHUNK; ++k)
              if (!PAIR_DBG(2)) tc_mma_i8(d_tmem, da0 + (uint64_t)((k * 32) >> 4), db0 + (uint64_t)((k * 32) >> 4), idesc, 1u);
            tc_commit(b_empty + stage);  // chunk reusable once both issuers' MMAs have read it
          }
          __syncwarp();
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (leader) tc_commit(a_empty + ab);  // this array may be refilled
        __syncwarp();
      }
      if (leader) tc_commit(tmem_full_bar + slot);
      __syncwarp();
    }
  } else {
    // ================================================================== epilogue warps (independent of each other)
    const int quarter = warp & 3;
    const int part = warp >> 2;                // 64-channel part of the CTA's 128
    const int j = lane >> 2;                   // column of the 8-column block
    const int ch_slot = 64 * part + PAIR_NCH * (lane & 3);   // first channel of the thread, within the accumulator slot
    const int ch0 = PAIR_N * nhalf + ch_slot;                 // ... and within the layer
    const bool fast = args.bounded != 0;
    EpiRegs<PAIR_NCH> K;
    epi_init(consts, ch0, magic_smem, K);
    if constexpr (!POOL) {
      if (!PAIR_DBG(8)) {
        // software-pipelined over the warp's tiles (epilogue16.cuh epi_pipeline); accumulator row 16*half + 8*s + j of
        // the quarter = pixel (2*quarter + half, j) of image s of the pair
        int band = band0, t = 0, acc_it = 0;
        auto next = [&](EpiTile& e) -> bool {
          if (t == PAIR_T) {
            t = 0;
            band += band_step;
          }
          if (band >= args.num_bands) return false;
          const uint32_t slot = acc_it % PAIR_SLOTS;
          const int64_t img0 = (int64_t)band * (2 * PAIR_T) + 2 * t;
          e.t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * PAIR_N + 64 * part;
          e.full_bar = tmem_full_bar + slot;
          e.empty_bar = tmem_empty_bar + slot;
          e.parity = (acc_it / PAIR_SLOTS) & 1;
          e.out = args.y + ((img0 * IMG + 2 * quarter) * IMG + j) * (int64_t)COUT + ch0;
          e.valid0 = img0 < args.n_img;
          e.valid1 = img0 + 1 < args.n_img;
          ++t;
          ++acc_it;
          return true;
        };
        epi_pipeline<CHECK>(K, consts, ch0, fast, args.zp_out, args.lo, (int64_t)IMG * IMG * COUT, (int64_t)IMG * COUT, lane,
                            next);
      }
    }
    if (POOL || PAIR_DBG(8)) {
      int acc_it = 0;
      for (int band = band0; band < args.num_bands; band += band_step) {
        for (int t = 0; t < PAIR_T; ++t, ++acc_it) {
          const uint32_t slot = acc_it % PAIR_SLOTS;
          const int64_t img0 = (int64_t)band * (2 * PAIR_T) + 2 * t;   // image 0 of the pair
          const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * PAIR_N + 64 * part;
          auto release = [&]() {
            if (lane == 0) mbar_arrive(tmem_empty_bar + slot);
          };
          mbar_wait(tmem_full_bar + slot, (acc_it / PAIR_SLOTS) & 1);
          tc_fence_after();
          if (PAIR_DBG(8)) {
            tc_fence_before();
            __syncwarp();
            release();
            continue;
          }
          if constexpr (POOL) {
            // thread (j, q): pooled pixel (row = quarter, column j>>1) of image j&1
            const int64_t img = img0 + (j & 1);
            uint8_t* out = args.y + ((img * (IMG / 2) + quarter) * (IMG / 2) + (j >> 1)) * (int64_t)COUT + ch0;
            epi_block_pool<CHECK, PAIR_NCH, /*PAIRED=*/true>(t_addr, K, consts, ch0, fast, args.zp_out, args.lo, out,
                                                            img < args.n_img, lane, release);
          }  // (!POOL reaches this loop only in the drain-only timing mode)
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PAIR_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem_base, PAIR_SLOTS * PAIR_N);
  }
}

template <int CIN, bool POOL, bool CHECK = true>
static int launch_pair(const uint8_t* x, uint8_t* y, int64_t n_img, const b200q_conv3x3* L, cudaStream_t stream) {
  using C = PairCfg<CIN>;
  const b200q_requant& rq = L->rq;
  if constexpr (CHECK) {  // drop the per-element range test when it is provably idle
    if ((rq.flags & B200Q_RQ_BOUNDED) && (rq.flags & B200Q_RQ_ACC22))
      return launch_pair<CIN, POOL, false>(x, y, n_img, L, stream);
  }
  CUtensorMap map_w;
  {
    // weights [COUT][9*CIN] viewed as (k, e, Q, b, u) with channel = 16*Q + 4*u + 2*b + e: a box of (KC, 2, 4, 2, 4)
    // lands as 64 rows in the order e + 2*q + 8*b + 16*u = the epilogue's column order (epi_channel_of_column<16>)
    const uint64_t ktot = 9ull * CIN;
    const uint64_t dims[5] = {ktot, 2, (uint64_t)PAIR_COUT / 16, 2, 4};
    const uint64_t strides[4] = {ktot, 16 * ktot, 2 * ktot, 4 * ktot};
    const uint32_t box[5] = {(uint32_t)C::KC, 2, 4, 2, 4};
    int rc = encode_tensor_map(&map_w, L->w, 5, dims, strides, box, C::KC);
    if (rc) return rc;
  }
  PairConsts consts;
  for (int c = 0; c < PAIR_COUT; ++c) {
    const int32_t corr = L->corr_host[4 * PAIR_COUT + c];  // class 4 = all nine taps (pads hold the zero-point)
    consts.cm[c] = (int32_t)(MAGIC_BITS - (uint32_t)corr);
    consts.k1[c] = -(MAGIC_F + (float)corr);  // exact: |corr| < 2^22 is part of B200Q_RQ_BOUNDED
    consts.bdiv[c] = rq.bdiv_host[c];
    consts.mult[c] = rq.mult_host[c];
  }
  auto kernel = conv_pair_kernel<CIN, POOL, CHECK>;
  static uint64_t attr_mask = 0;  // per template instantiation
  if (int arc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), C::SMEM_BYTES, &attr_mask)) return arc;
  const int num_bands = (int)((n_img + 2 * PAIR_T - 1) / (2 * PAIR_T));
#ifdef B200Q_DEV
  static int debug = -1;
  if (debug < 0) {
    const char* e = getenv("B200Q_PAIR_DEBUG");
    debug = e ? atoi(e) : 0;
  }
#else
  const int debug = 0;
#endif
  PairArgs args{x,         y,
                n_img,     num_bands,
                L->zp_x,   rq.zp_out,
                rq.relu ? rq.zp_out : 0, (rq.flags & B200Q_RQ_BOUNDED) ? 1 : 0,
                debug};
  // two CTAs (one per channel half) per band; an even grid no larger than the SM count
  int grid = 2 * num_bands < num_sms() ? 2 * num_bands : (num_sms() & ~1);
  return launch_kernel("conv_pair_kernel", kernel, grid, PAIR_THREADS, C::SMEM_BYTES, stream, map_w, consts, args);
}

// Entry used by b200q_conv3x3_tc for the geometries this kernel covers; returns 1 when the geometry is not handled.
int conv3x3_pair_dispatch(const uint8_t* x, uint8_t* y, int64_t b, const b200q_conv3x3* L, bool pool, cudaStream_t s,
                          int* rc) {
  if (!L->corr_host || !L->rq.mult_host || !L->rq.bdiv_host) return 1;
  if (L->img != PAIR_IMG || L->cout != PAIR_COUT) return 1;
  if (L->cin == 128) {
    *rc = pool ? launch_pair<128, true>(x, y, b, L, s) : launch_pair<128, false>(x, y, b, L, s);
    return 0;
  }
  if (L->cin == 256) {
    *rc = pool ? launch_pair<256, true>(x, y, b, L, s) : launch_pair<256, false>(x, y, b, L, s);
    return 0;
  }
  return 1;
}

}  // namespace b200q
