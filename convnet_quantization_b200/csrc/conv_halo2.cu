// conv2 (3x3, 64 -> 64 channels on 32x32 images, + ReLU + fused 2x2 max-pool) on a CTA PAIR: tcgen05.mma.cta_group::2.
//
// conv_halo.cu runs this layer at the shared-memory operand limit, not the tensor-pipe limit: an M=128, N=64, K=32 MMA
// reads 4 KB of activations + 2 KB of weights at 128 B/clk = 48 cycles against 32 cycles of tensor work (67 % cap; 64 %
// achieved).  N is the layer's 64 output channels and cannot grow, but the WEIGHT operand can be shared: two CTAs on the
// SMs of one TPC form a cluster, each keeps one image band (as in conv_halo.cu) and HALF of the weight rows, and one
// thread of the leader CTA issues M=256 MMAs that cover tile t of BOTH images.  Per CTA and MMA: 4 KB of A + 1 KB of
// B = 40 cycles.
//
// Everything else is conv_halo.cu's design (padded pixel-sequence band, nine row-shifted views of one array, 8x16-pixel
// block tiles, pre-biased accumulators, fragment-layout epilogue with the pool fused in front of the requantisation),
// plus the pair protocol:
//   * barriers that only the leader's MMA issuers wait on live in the LEADER's shared memory: `peer_full` (the peer's band
//     has landed: forwarded by the peer's otherwise idle issuer warp after its own proxy fence) and `tmem_empty` (both
//     CTAs' epilogue warps arrive on it, the peer's through mapa / shared::cluster);
//   * MMA completion is multicast (`tcgen05.commit...multicast::cluster`, mask 0b11) to `tmem_full` and to the band's
//     `empty` barrier at the same shared-memory offset in both CTAs;
//   * TMEM is allocated with cta_group::2 by one warp in each CTA and freed after a cluster barrier.
// Pair p of the grid handles images 2p and 2p+1; an odd last image leaves the peer CTA computing on stale data that is
// never stored.
#include "common.cuh"
#include "epilogue16.cuh"

namespace b200q {

namespace h2 {
constexpr int IMG = 32, CIN = 64, COUT = 64;
constexpr int EW = 8, ISSUERS = 2;
constexpr int LOAD_WARP = EW, MMA_WARP = EW + 1;
constexpr int THREADS = 32 * (EW + 1 + ISSUERS);
constexpr int P = IMG + 1;
constexpr int POS_PER_IMG = (IMG + 1) * P;
constexpr int A_POS = POS_PER_IMG + P + 2;
constexpr int A_BYTES = (A_POS * CIN + 1023) / 1024 * 1024;
constexpr int TILES_X = IMG / 8, TILES_Y = IMG / 16, TILES = TILES_X * TILES_Y;  // 8 tiles of 8 columns x 16 rows
constexpr int W_ROWS = COUT / 2;                 // weight rows (output channels) held by each CTA of the pair
constexpr int W_TAP_BYTES = W_ROWS * CIN;        // 2 KB
constexpr int W_BYTES = 9 * W_TAP_BYTES;
constexpr int SMEM_BYTES = 2 * A_BYTES + W_BYTES + 256 + 1024;
constexpr int NCH = 16, SETS = 2;                // epilogue: 8 warps x 16 channels per thread, two sets of four warps
constexpr int MAX_SLOTS = 8;                     // accumulator slots: 4 or 8 x 64 columns of TMEM (template parameter)
static_assert(TILES % ISSUERS == 0, "tiles per issuer");
}  // namespace h2

struct Halo2Args {
  const uint8_t* x;
  uint8_t* y;
  const int8_t* w;  // [COUT][9][CIN]
  int64_t n_img;
  int num_pairs;
  int zp_x, zp_out, lo, bounded;
};

struct alignas(16) Halo2Consts {
  int32_t cm[h2::COUT];
  float k1[h2::COUT];
  float bdiv[h2::COUT];
  float mult[h2::COUT];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] += A[each CTA's own 128 rows] * B[each CTA's half of the N rows]; issued by one thread of the leader
__device__ __forceinline__ void tc_mma_i8_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of this thread's cta_group::2 MMAs -> arrive on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_2cta(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

template <bool CHECK, int SLOTS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(h2::THREADS, 1)
conv_halo2_kernel(const __grid_constant__ Halo2Consts consts, const Halo2Args args) {
  using namespace h2;
  constexpr int TMEM_COLS = SLOTS * COUT;
  static_assert(SLOTS % SETS == 0 && SLOTS <= MAX_SLOTS && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "slots");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_smem = smem;                                  // 2 x A_BYTES
  uint8_t* w_smem = a_smem + 2 * A_BYTES;                  // 9 x [W_ROWS][64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(w_smem + W_BYTES);  // [2] this CTA's band landed
  uint64_t* peer_full_bar = full_bar + 2;                  // [2] leader only: the peer's band landed
  uint64_t* empty_bar = peer_full_bar + 2;                 // [2] band consumed (multicast commit)
  uint64_t* w_bar = empty_bar + 2;                         // weights landed
  uint64_t* tmem_full_bar = w_bar + 1;                     // [SLOTS] multicast commit
  uint64_t* tmem_empty_bar = tmem_full_bar + MAX_SLOTS;    // [SLOTS] leader only: drained by both CTAs
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + MAX_SLOTS);
  uint32_t* magic_smem = tmem_base_smem + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)
  pdl_launch_dependents();

  if (warp == LOAD_WARP && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(full_bar + i, 32);
      mbar_init(peer_full_bar + i, 1);
      mbar_init(empty_bar + i, ISSUERS);
    }
    for (int i = 0; i < SLOTS; ++i) {
      mbar_init(tmem_full_bar + i, 1);
      mbar_init(tmem_empty_bar + i, 2 * 4);  // four warps of one set, in each of the two CTAs
    }
    mbar_init(w_bar, 32 * (EW + 1));
    *magic_smem = MAGIC_BITS;
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc_2cta(tmem_base_smem, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  if (warp < EW) {  // pad positions hold the activation zero-point for the lifetime of the CTA (see conv_halo.cu)
    const int t = threadIdx.x;
    const uint32_t zp4 = (uint32_t)args.zp_x * 0x01010101u;
    const uint4 zpv = make_uint4(zp4, zp4, zp4, zp4);
    for (int buf = 0; buf < 2; ++buf) {
      uint8_t* a_buf = a_smem + buf * A_BYTES;
      uint4* tail = reinterpret_cast<uint4*>(a_buf + POS_PER_IMG * CIN);
      for (int i = t; i < (A_BYTES - POS_PER_IMG * CIN) / 16; i += 32 * EW) tail[i] = zpv;
      constexpr int PADS = P + IMG;
      for (int i = t; i < PADS * (CIN / 16); i += 32 * EW) {
        const int pad = i / (CIN / 16), part = i % (CIN / 16);
        const int pos = pad < P ? pad : (pad - P + 1) * P;
        *reinterpret_cast<uint4*>(a_buf + pos * CIN + part * 16) = zpv;
      }
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  if (warp < 4) {  // pre-bias every accumulator slot of THIS CTA's TMEM (see requant4_prebiased)
    const uint32_t base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int slot = 0; slot < SLOTS; ++slot)
      for (int c = 0; c < COUT; c += 8) tmem_st_fill8(base + slot * COUT + c, MAGIC_BITS);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp <= LOAD_WARP) {
    // This CTA's half of the weights: accumulator columns (= B rows) 32*rank .. 32*rank+31 of every tap, i.e. output
    // channels epi_channel_of_column<16>(32*rank + n), as [32][64 B] SW64 K-major blocks of 2 KB per tap.
    constexpr int CPR = CIN / 16;
    const uint32_t w_base = smem_u32(w_smem);
    for (int g = threadIdx.x; g < 9 * W_ROWS * CPR; g += 32 * (EW + 1)) {
      const int part = g % CPR, n = (g / CPR) % W_ROWS, tap = g / (CPR * W_ROWS);
      const int swz = (n >> 1) & 3;
      const uint32_t dst = w_base + tap * W_TAP_BYTES + n * CIN + ((part ^ swz) << 4);
      const int8_t* src = args.w + ((int64_t)epi_channel_of_column<NCH>(W_ROWS * (int)rank + n) * 9 + tap) * CIN + part * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(w_bar)) : "memory");
  }

  const int pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;

  if (warp == LOAD_WARP) {
    // ================================================================== loader: this CTA's image of the pair
    constexpr int CHUNKS_PER_IMG = IMG * IMG * (CIN / 16);
    pdl_wait();
    int it = 0;
    for (int pair = pair0; pair < args.num_pairs; pair += pair_step, ++it) {
      const int buf = it & 1;
      mbar_wait(empty_bar + buf, ((it >> 1) & 1) ^ 1);
      const uint32_t a_buf = smem_u32(a_smem + buf * A_BYTES);
      const int64_t img = 2 * (int64_t)pair + rank;
      if (img < args.n_img) {
        const uint8_t* src = args.x + img * (int64_t)(IMG * IMG * CIN);
#pragma unroll 8
        for (int g = lane; g < CHUNKS_PER_IMG; g += 32) {
          const int px = g / (CIN / 16), part = g % (CIN / 16);
          const int pos = (px / IMG + 1) * P + (px % IMG) + 1;
          const int swz = (pos >> 1) & 3;
          const uint32_t dst = a_buf + pos * CIN + ((part ^ swz) << 4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + g * 16) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(full_bar + buf)) : "memory");
    }
  } else if (warp >= MMA_WARP) {
    const int issuer = warp - MMA_WARP;
    if (rank != 0) {
      // ================================================================== peer: forward "my band has landed" to the leader
      if (issuer == 0) {
        mbar_wait(w_bar, 0);
        int it = 0;
        for (int pair = pair0; pair < args.num_pairs; pair += pair_step, ++it) {
          const int buf = it & 1;
          mbar_wait(full_bar + buf, (it >> 1) & 1);
          fence_proxy_async_smem();  // this CTA's cp.async data -> visible to the (leader-issued) tensor-core reads
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(peer_full_bar + buf, 0);
        }
      }
    } else {
      // ================================================================== leader: MMA issuers (alternate tiles)
      const bool leader = elect_one() != 0;
      constexpr uint32_t idesc = make_idesc_i8(256, COUT);
      mbar_wait(w_bar, 0);
      fence_proxy_async_smem();
      const uint64_t w_desc0 = make_kmajor_desc<CIN>(smem_u32(w_smem), 8 * CIN);
      int it = 0;
      for (int pair = pair0; pair < args.num_pairs; pair += pair_step, ++it) {
        const int buf = it & 1;
        mbar_wait(full_bar + buf, (it >> 1) & 1);
        mbar_wait(peer_full_bar + buf, (it >> 1) & 1);
        fence_proxy_async_smem();
        tc_fence_after();
        const uint64_t a_desc0 = make_kmajor_desc<CIN>(smem_u32(a_smem + buf * A_BYTES), P * CIN);
        for (int t = issuer; t < TILES; t += ISSUERS) {
          const int acc_it = it * TILES + t;
          const uint32_t slot = acc_it % SLOTS;
          mbar_wait(tmem_empty_bar + slot, ((acc_it / SLOTS) & 1) ^ 1);
          tc_fence_after();
          if (leader) {
            const int r0 = (t / TILES_X) * 16, c0 = (t % TILES_X) * 8;
            const uint32_t d_tmem = tmem_base + slot * COUT;
            const uint64_t a_tile = a_desc0 + (uint64_t)(((r0 * P + c0) * CIN) >> 4);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int k = 0; k < CIN / 32; ++k) {
                const uint64_t da = a_tile + (uint64_t)((((tap / 3) * P + (tap % 3)) * CIN + k * 32) >> 4);
                const uint64_t db = w_desc0 + (uint64_t)((tap * W_TAP_BYTES + k * 32) >> 4);
                tc_mma_i8_2cta(d_tmem, da, db, idesc, 1u);
              }
            }
            tc_commit_2cta(tmem_full_bar + slot);
          }
          __syncwarp();
        }
        if (leader) tc_commit_2cta(empty_bar + buf);  // both CTAs' bands are free once these MMAs have read them
        __syncwarp();
      }
    }
  } else {
    // ================================================================== epilogue warps (each CTA drains its own TMEM)
    const int quarter = warp & 3;
    const int set = warp >> 2;
    const int j = lane >> 2;
    const int ch0 = NCH * (lane & 3);
    const bool fast = args.bounded != 0;
    EpiRegs<NCH> K;
    epi_init(consts, ch0, magic_smem, K);
    int acc_base = 0;
    for (int pair = pair0; pair < args.num_pairs; pair += pair_step, acc_base += TILES) {
      const int64_t img = 2 * (int64_t)pair + rank;
      const bool valid = img < args.n_img;
      for (int t = (set - acc_base % SETS + SETS) % SETS; t < TILES; t += SETS) {
        const int acc_it = acc_base + t;
        const uint32_t slot = acc_it % SLOTS;
        const int r0 = (t / TILES_X) * 16 + 4 * quarter, c = (t % TILES_X) * 8 + j;
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * COUT;
        auto release = [&]() {
          if (lane == 0) mbar_arrive_cluster(tmem_empty_bar + slot, 0);  // the leader's barrier counts both CTAs
        };
        mbar_wait(tmem_full_bar + slot, (acc_it / SLOTS) & 1);
        tc_fence_after();
        uint8_t* out = args.y + ((img * (IMG / 2) + (r0 >> 1) + (j & 1)) * (IMG / 2) + (c >> 1)) * (int64_t)COUT + ch0;
        epi_block_pool<CHECK>(t_addr, K, consts, ch0, fast, args.zp_out, args.lo, out, valid, lane, release);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's tensor-core reads of this CTA's shared memory / TMEM are over
  if (warp == MMA_WARP) {
    __syncwarp();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

// Entry used by b200q_conv3x3_tc for conv2 + fused pool; returns 1 when the layer / batch is not covered.
int conv3x3_halo2_dispatch(const uint8_t* x, uint8_t* y, int64_t b, const b200q_conv3x3* L, bool pool, cudaStream_t s,
                           int* rc) {
  using namespace h2;
  const b200q_requant& rq = L->rq;
  if (!pool || L->img != IMG || L->cin != CIN || L->cout != COUT) return 1;
  if (!L->corr_host || !rq.mult_host || !rq.bdiv_host) return 1;
  Halo2Consts consts;
  for (int c = 0; c < COUT; ++c) {
    const int32_t corr = L->corr_host[4 * COUT + c];
    consts.cm[c] = (int32_t)(MAGIC_BITS - (uint32_t)corr);
    consts.k1[c] = -(MAGIC_F + (float)corr);
    consts.bdiv[c] = rq.bdiv_host[c];
    consts.mult[c] = rq.mult_host[c];
  }
  const bool check = !((rq.flags & B200Q_RQ_BOUNDED) && (rq.flags & B200Q_RQ_ACC22));
  // Eight accumulator slots (all 512 TMEM columns): a slot returns to the leader only when BOTH CTAs' epilogue warps have
  // drained it, so the issuers need more look-ahead than in the single-CTA kernel.  Development builds: B200Q_H2_SLOTS=4.
  int slots = 8;
#ifdef B200Q_DEV
  {
    static int env = -1;
    if (env < 0) {
      const char* e = getenv("B200Q_H2_SLOTS");
      env = e ? atoi(e) : 0;
    }
    if (env == 4) slots = 4;
  }
#endif
  void (*kernel)(Halo2Consts, Halo2Args) =
      slots == 8 ? (check ? conv_halo2_kernel<true, 8> : conv_halo2_kernel<false, 8>)
                 : (check ? conv_halo2_kernel<true, 4> : conv_halo2_kernel<false, 4>);
  static uint64_t attr_mask[4] = {0, 0, 0, 0};
  *rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), SMEM_BYTES, &attr_mask[(check ? 1 : 0) + (slots == 8 ? 2 : 0)]);
  if (*rc) return 0;
  const int num_pairs = (int)((b + 1) / 2);
  Halo2Args args{x, y, L->w, b, num_pairs, L->zp_x, rq.zp_out, rq.relu ? rq.zp_out : 0, (rq.flags & B200Q_RQ_BOUNDED) ? 1 : 0};
  // persistent: as many CTA pairs as the device can hold at once (a pair needs both SMs of one TPC)
  static int max_clusters_cached[64] = {0};
  int dev = 0;
  if (int drc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) {
    *rc = drc;
    return 0;
  }
  int max_pairs = (dev >= 0 && dev < 64) ? max_clusters_cached[dev] : 0;  // (same for every instantiation: same resources)
  if (max_pairs == 0) {
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3((unsigned)(num_sms() & ~1));
    q.blockDim = dim3(THREADS);
    q.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2;
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    q.attrs = qa;
    q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &q) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      return 1;  // no cluster support reported: let the single-CTA kernel take the layer
    }
    max_pairs = n < num_sms() / 2 ? n : num_sms() / 2;
    if (dev >= 0 && dev < 64) max_clusters_cached[dev] = max_pairs;
  }
  const int grid = 2 * (num_pairs < max_pairs ? num_pairs : max_pairs);
  *rc = launch_kernel("conv_halo2_kernel", kernel, grid, THREADS, SMEM_BYTES, s, consts, args);  // cluster dims: __cluster_dims__
  return 0;
}

}  // namespace b200q
