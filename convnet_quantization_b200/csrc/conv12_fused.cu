// conv1 and conv2 in ONE kernel:  fp32 NCHW [b,3,32,32] -> quantize -> conv1+ReLU -> conv2+ReLU -> 2x2 max-pool ->
// uint8 NHWC [b,16,16,64].
//
// Run separately (conv1_tc.cu, conv_halo.cu) the two layers have complementary bottlenecks: conv1 is bound by the
// instruction issue of its epilogue warps (one MMA per tile, 65 536 requantisations per image; tensor pipe 4 % busy),
// conv2 by the tensor pipe's shared-memory operand fetch (144 MMAs per image; issue slots mostly idle).  Here one CTA
// per SM does both, one image at a time: the conv1 epilogue writes its uint8 output straight into conv2's padded
// activation band in shared memory (never to HBM: -128 KB of traffic per image), and while the tensor pipe works
// through conv2 of image i the CUDA cores quantise, gather and requantise conv1 of image i+1.
//
// Roles (14 warps = 448 threads, 128 registers per thread):
//   warps 0..7   epilogue of BOTH layers (epilogue16.cuh low-register variants).  Per image slot a warp alternates
//                conv1 blocks of image i+1 (store: shared-memory band) and conv2 blocks of image i (2x2 max-pool, store:
//                global), swapping its 48 requantisation constants from shared memory in between; two sets of four
//                warps take alternate tiles of either layer.
//   warps 8..11  conv1 producers (as conv1_tc.cu: exact quantisation, im2col rows; warp 8 issues conv1's 8 MMAs).
//   warps 12,13  conv2 MMA issuers (alternate tiles, as conv_halo.cu); they also copy conv2's weights once.
// TMEM: columns [0,256) = four conv1 accumulator slots, [256,512) = four conv2 slots, all pre-biased.
// Shared memory: two conv2 bands (2 x 70 KB), conv2 weights (36 KB), one conv1 im2col buffer (32 KB), conv1 weights,
// the quantised image, both constant tables: 218 KB.
#include "common.cuh"
#include "epilogue16.cuh"

namespace b200q {

namespace f12 {
constexpr int EPI_WARPS = 8, PROD_WARPS = 4, ISSUERS = 2;
constexpr int PROD_WARP0 = EPI_WARPS, MMA_WARP0 = EPI_WARPS + PROD_WARPS;
constexpr int THREADS = 32 * (EPI_WARPS + PROD_WARPS + ISSUERS);
constexpr int IMG = 32, COUT = 64, TILES = 8, SLOTS = 4;
// conv1 (as conv1_tc.cu)
constexpr int KB1 = 32;                                   // conv1 K bytes per im2col row
constexpr int A1_TILE = 128 * KB1, A1_BYTES = TILES * A1_TILE;   // 32 KB, single buffer
constexpr int B1_BYTES = COUT * KB1;
constexpr int QP = IMG + 2, Q_BYTES = (QP * QP * 4 + 15) / 16 * 16;
// conv2 (as conv_halo.cu<32,64,64,1>)
constexpr int CIN2 = 64, P2 = IMG + 1, POS2 = (IMG + 1) * P2, A2_POS = POS2 + P2 + 2;
constexpr int A2_BYTES = (A2_POS * CIN2 + 1023) / 1024 * 1024;  // 70 KB per band
constexpr int W2_TAP = COUT * CIN2, W2_BYTES = 9 * W2_TAP;
constexpr int CONST_FLOATS = 2 * 3 * COUT;                 // [layer][k1|bdiv|mult][channel]
constexpr int SMEM = 2 * A2_BYTES + W2_BYTES + A1_BYTES + B1_BYTES + Q_BYTES + CONST_FLOATS * 4 + 512 + 1024;
static_assert(SMEM <= 227 * 1024, "shared memory");
}  // namespace f12

struct F12Consts {  // per output channel, both layers (kernel parameter; also the source of the shared-memory tables)
  int32_t cm[2][f12::COUT];
  float k1[2][f12::COUT];
  float bdiv[2][f12::COUT];
  float mult[2][f12::COUT];
};
// view of one layer with the member names the epilogue's exact fall-back expects
struct F12LayerConsts {
  const int32_t* cm;
  const float *k1, *bdiv, *mult;
};

struct F12Args {
  const float* x;
  uint8_t* y;
  const int8_t* w1;    // [64][9][4]
  const int8_t* w2;    // [64][9][64]
  int64_t n_img;
  float inv_scale;
  int zp_in;           // zero-point of the quantised input (conv1 pads)
  int zp1, lo1;        // conv1 output zero-point / lower clamp  (= conv2 input zero-point: conv2 pads)
  int zp2, lo2;
  int bounded1, bounded2;
};

__device__ __forceinline__ uint32_t f12_quantize_magic(float x, float inv_scale, int zp_sub) {
  float t = __fmul_rn(x, inv_scale);
  t = fminf(fmaxf(t, -1024.0f), 1024.0f);
  const int q = __float_as_int(__fadd_rn(t, MAGIC_F)) + zp_sub;
  return (uint32_t)max(0, min(q, 255));
}

template <bool CHECK1, bool CHECK2>
__global__ void __launch_bounds__(f12::THREADS, 1)
conv12_fused_kernel(const __grid_constant__ F12Consts consts, const F12Args args) {
  using namespace f12;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a2_smem = smem;                                  // [2][A2_BYTES] conv2 bands
  uint8_t* w2_smem = a2_smem + 2 * A2_BYTES;                // 9 x [64][64]
  uint8_t* a1_smem = w2_smem + W2_BYTES;                    // [8 tiles][128 rows][32 B]
  uint8_t* b1_smem = a1_smem + A1_BYTES;                    // [64][32 B]
  uint32_t* q_img = reinterpret_cast<uint32_t*>(b1_smem + B1_BYTES);
  float* c_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(q_img) + Q_BYTES);   // [2][3][64]
  uint64_t* a1_empty = reinterpret_cast<uint64_t*>(c_smem + CONST_FLOATS);  // [2] half of the im2col buffer consumed
  uint64_t* band_full = a1_empty + 2;                       // [2] conv1 output of an image complete in its band
  uint64_t* band_empty = band_full + 2;                     // [2] conv2 MMAs of the band's image complete
  uint64_t* w2_bar = band_empty + 2;
  uint64_t* t1_full = w2_bar + 1;                           // [SLOTS] conv1 accumulator slots
  uint64_t* t1_empty = t1_full + SLOTS;
  uint64_t* t2_full = t1_empty + SLOTS;                     // [SLOTS] conv2 accumulator slots
  uint64_t* t2_empty = t2_full + SLOTS;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(t2_empty + SLOTS);
  uint32_t* magic_smem = tmem_base_smem + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == PROD_WARP0 && lane == 0) {
    *magic_smem = MAGIC_BITS;
    mbar_init(a1_empty, 1);
    mbar_init(a1_empty + 1, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(band_full + i, EPI_WARPS);
      mbar_init(band_empty + i, ISSUERS);
    }
    mbar_init(w2_bar, 32 * ISSUERS);
    for (int i = 0; i < SLOTS; ++i) {
      mbar_init(t1_full + i, 1);
      mbar_init(t1_empty + i, 4);
      mbar_init(t2_full + i, 1);
      mbar_init(t2_empty + i, 4);
    }
    fence_barrier_init();
  }
  if (warp == PROD_WARP0) {
    tmem_alloc(tmem_base_smem, 2 * SLOTS * COUT);
    tmem_relinquish();
  }
  if (warp < EPI_WARPS) {
    const int t = threadIdx.x;
    // conv1: border of the quantised image = input zero-point
    const uint32_t zpi4 = (uint32_t)args.zp_in * 0x01010101u;
    for (int i = t; i < QP * QP; i += 32 * EPI_WARPS) {
      const int r = i / QP, c = i % QP;
      if (r == 0 || r == QP - 1 || c == 0 || c == QP - 1) q_img[i] = zpi4;
    }
    // conv1 weights [64][9][4] -> B operand [64][k = tap*3 + ch] (32-byte rows, SWIZZLE_32B), rows in epilogue order
    for (int i = t; i < COUT * (KB1 / 4); i += 32 * EPI_WARPS) {
      const int nr = i / (KB1 / 4), wd = i % (KB1 / 4);
      const int n = epi_channel_of_column<16>(nr);
      uint32_t word = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int k = wd * 4 + b;
        if (k < 27) word |= (uint32_t)(uint8_t)__ldg(args.w1 + (n * 9 + k / 3) * 4 + k % 3) << (8 * b);
      }
      const int chunk = (wd >> 2) ^ ((nr >> 2) & 1);
      *reinterpret_cast<uint32_t*>(b1_smem + nr * KB1 + chunk * 16 + (wd & 3) * 4) = word;
    }
    // conv2 bands: pad positions (row -1, column -1 of the image, everything behind it) = conv1's output zero-point
    const uint32_t zp4 = (uint32_t)args.zp1 * 0x01010101u;
    const uint4 zpv = make_uint4(zp4, zp4, zp4, zp4);
    for (int buf = 0; buf < 2; ++buf) {
      uint8_t* a_buf = a2_smem + buf * A2_BYTES;
      uint4* tail = reinterpret_cast<uint4*>(a_buf + POS2 * CIN2);
      for (int i = t; i < (A2_BYTES - POS2 * CIN2) / 16; i += 32 * EPI_WARPS) tail[i] = zpv;
      constexpr int PADS = P2 + IMG;
      for (int i = t; i < PADS * (CIN2 / 16); i += 32 * EPI_WARPS) {
        const int pad = i / (CIN2 / 16), part = i % (CIN2 / 16);
        const int pos = pad < P2 ? pad : (pad - P2 + 1) * P2;
        *reinterpret_cast<uint4*>(a_buf + pos * CIN2 + part * 16) = zpv;
      }
    }
    // requantisation constants of both layers
    for (int i = t; i < 2 * COUT; i += 32 * EPI_WARPS) {
      const int l = i / COUT, c = i % COUT;
      c_smem[(l * 3 + 0) * COUT + c] = consts.k1[l][c];
      c_smem[(l * 3 + 1) * COUT + c] = consts.bdiv[l][c];
      c_smem[(l * 3 + 2) * COUT + c] = consts.mult[l][c];
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  if (warp < 4) {  // pre-bias all eight accumulator slots
    const uint32_t base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 2 * SLOTS * COUT; c += 8) tmem_st_fill8(base + c, MAGIC_BITS);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int my_imgs = ((int64_t)blockIdx.x < args.n_img) ? (int)((args.n_img - 1 - blockIdx.x) / gridDim.x + 1) : 0;

  if (warp >= MMA_WARP0) {
    // ================================================================== conv2 MMA issuers (alternate tiles)
    const int issuer = warp - MMA_WARP0;
    {  // conv2 weights, once: global [64][9][64] -> nine [64][64] K-major swizzled tap blocks, rows in epilogue order
      constexpr int CPR = CIN2 / 16;
      const uint32_t w_base = smem_u32(w2_smem);
      for (int g = issuer * 32 + lane; g < 9 * COUT * CPR; g += 32 * ISSUERS) {
        const int part = g % CPR, n = (g / CPR) % COUT, tap = g / (CPR * COUT);
        const int swz = (n >> 1) & 3;
        const uint32_t dst = w_base + tap * W2_TAP + n * CIN2 + ((part ^ swz) << 4);
        const int8_t* src = args.w2 + ((int64_t)epi_channel_of_column<16>(n) * 9 + tap) * CIN2 + part * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(w2_bar)) : "memory");
    }
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc_i8(128, COUT);
    mbar_wait(w2_bar, 0);
    fence_proxy_async_smem();
    const uint64_t w_desc0 = make_kmajor_desc<CIN2>(smem_u32(w2_smem), 8 * CIN2);
    for (int it = 0; it < my_imgs; ++it) {
      const int buf = it & 1;
      mbar_wait(band_full + buf, (it >> 1) & 1);
      fence_proxy_async_smem();
      tc_fence_after();
      const uint64_t a_desc0 = make_kmajor_desc<CIN2>(smem_u32(a2_smem + buf * A2_BYTES), P2 * CIN2);
      for (int t = issuer; t < TILES; t += ISSUERS) {
        const int acc_it = it * TILES + t;
        const uint32_t slot = acc_it % SLOTS;
        mbar_wait(t2_empty + slot, ((acc_it / SLOTS) & 1) ^ 1);
        tc_fence_after();
        if (leader) {
          const int r0 = (t / 4) * 16, c0 = (t % 4) * 8;
          const uint32_t d_tmem = tmem_base + (SLOTS + slot) * COUT;
          const uint64_t a_tile = a_desc0 + (uint64_t)(((r0 * P2 + c0) * CIN2) >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int k = 0; k < CIN2 / 32; ++k) {
              const uint64_t da = a_tile + (uint64_t)((((tap / 3) * P2 + (tap % 3)) * CIN2 + k * 32) >> 4);
              const uint64_t db = w_desc0 + (uint64_t)((tap * W2_TAP + k * 32) >> 4);
              tc_mma_i8(d_tmem, da, db, idesc, 1u);
            }
          }
          tc_commit(t2_full + slot);
        }
        __syncwarp();
      }
      if (leader) tc_commit(band_empty + buf);
      __syncwarp();
    }
  } else if (warp >= PROD_WARP0) {
    // ================================================================== conv1 producers (see conv1_tc.cu)
    const int p = threadIdx.x - 32 * PROD_WARP0;
    const int zp_sub = args.zp_in - (int)MAGIC_BITS;
    const uint32_t zp_hi = (uint32_t)args.zp_in << 24;
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc_i8(128, COUT);
    const uint64_t b_desc = make_kmajor_desc<KB1>(smem_u32(b1_smem), 8 * KB1);
    const uint64_t a_desc0 = make_kmajor_desc<KB1>(smem_u32(a1_smem), 8 * KB1);
    const int row = p >> 2, col0 = (p & 3) * 8;
    float4 v[3][2];
    auto load_image = [&](int it) {
      const int64_t img = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      const float* src = args.x + img * (3 * IMG * IMG) + row * IMG + col0;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        v[ch][0] = __ldg(reinterpret_cast<const float4*>(src + ch * IMG * IMG));
        v[ch][1] = __ldg(reinterpret_cast<const float4*>(src + ch * IMG * IMG) + 1);
      }
    };
    if (my_imgs > 0) load_image(0);
    for (int it = 0; it < my_imgs; ++it) {
      {  // (1) quantise
        uint32_t* dst = q_img + (row + 1) * QP + col0 + 1;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float c0[4] = {v[0][h].x, v[0][h].y, v[0][h].z, v[0][h].w};
          const float c1[4] = {v[1][h].x, v[1][h].y, v[1][h].z, v[1][h].w};
          const float c2[4] = {v[2][h].x, v[2][h].y, v[2][h].z, v[2][h].w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[h * 4 + j] = f12_quantize_magic(c0[j], args.inv_scale, zp_sub) |
                             (f12_quantize_magic(c1[j], args.inv_scale, zp_sub) << 8) |
                             (f12_quantize_magic(c2[j], args.inv_scale, zp_sub) << 16) | zp_hi;
        }
      }
      if (it + 1 < my_imgs) load_image(it + 1);
      asm volatile("bar.sync 2, %0;" ::"n"(32 * PROD_WARPS) : "memory");
      // (2)+(3) per half image (tiles 0-3, tiles 4-7): im2col rows into that half of the single buffer - free once the
      // previous image's MMAs of the same half have read it - then warp 8 issues the half's four MMAs.  With the
      // barrier per half the producers are never on the critical path: a half's rows are written while the epilogue
      // still drains the other half's accumulator slots.
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        mbar_wait(a1_empty + hf, (it & 1) ^ 1);
#pragma unroll 2
        for (int i = 4 * hf; i < 4 * hf + 4; ++i) {
          const int r = (i >> 2) * 16 + (p >> 3), c = (i & 3) * 8 + (p & 7);
          const uint32_t* q = q_img + r * QP + c;
          const uint32_t s0 = q[0], s1 = q[1], s2 = q[2];
          const uint32_t s3 = q[QP], s4 = q[QP + 1], s5 = q[QP + 2];
          const uint32_t s6 = q[2 * QP], s7 = q[2 * QP + 1], s8 = q[2 * QP + 2];
          const uint4 lo4 = make_uint4(__byte_perm(s0, s1, 0x4210), __byte_perm(s1, s2, 0x5421),
                                       __byte_perm(s2, s3, 0x6542), __byte_perm(s4, s5, 0x4210));
          const uint4 hi4 = make_uint4(__byte_perm(s5, s6, 0x5421), __byte_perm(s6, s7, 0x6542), s8 & 0x00ffffffu, 0u);
          const int sw = (p >> 2) & 1;
          uint8_t* rowp = a1_smem + i * A1_TILE + p * KB1;
          *reinterpret_cast<uint4*>(rowp + (sw << 4)) = lo4;
          *reinterpret_cast<uint4*>(rowp + ((sw ^ 1) << 4)) = hi4;
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 2, %0;" ::"n"(32 * PROD_WARPS) : "memory");
        if (warp == PROD_WARP0) {
          tc_fence_after();
          for (int t = 4 * hf; t < 4 * hf + 4; ++t) {
            const int acc_it = it * TILES + t;
            const uint32_t slot = acc_it % SLOTS;
            mbar_wait(t1_empty + slot, ((acc_it / SLOTS) & 1) ^ 1);
            tc_fence_after();
            if (leader) {
              tc_mma_i8(tmem_base + slot * COUT, a_desc0 + (uint64_t)((t * A1_TILE) >> 4), b_desc, idesc, 1u);
              tc_commit(t1_full + slot);
            }
            __syncwarp();
          }
          if (leader) tc_commit(a1_empty + hf);
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================================== epilogue warps, both layers
    const int quarter = warp & 3;
    const int set = warp >> 2;                 // tiles t with t % 2 == set, of either layer
    const int j = lane >> 2;
    const int q4 = lane & 3;
    const int ch0 = 16 * q4;
    const bool fast1 = args.bounded1 != 0, fast2 = args.bounded2 != 0;
    const F12LayerConsts L1{consts.cm[0], consts.k1[0], consts.bdiv[0], consts.mult[0]};
    const F12LayerConsts L2{consts.cm[1], consts.k1[1], consts.bdiv[1], consts.mult[1]};
    EpiRegs<16> K;
    epi_init<16, /*PIN=*/false>(L1, ch0, magic_smem, K);
    auto load_consts = [&](int layer) {  // 48 constants of this thread's channels from the shared-memory tables
      const float4* k1 = reinterpret_cast<const float4*>(c_smem + (layer * 3 + 0) * COUT + ch0);
      const float4* bd = reinterpret_cast<const float4*>(c_smem + (layer * 3 + 1) * COUT + ch0);
      const float4* mu = reinterpret_cast<const float4*>(c_smem + (layer * 3 + 2) * COUT + ch0);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 a = k1[g], b = bd[g], c = mu[g];
        K.k1[4 * g] = a.x; K.k1[4 * g + 1] = a.y; K.k1[4 * g + 2] = a.z; K.k1[4 * g + 3] = a.w;
        K.bd[4 * g] = b.x; K.bd[4 * g + 1] = b.y; K.bd[4 * g + 2] = b.z; K.bd[4 * g + 3] = b.w;
        K.mu[4 * g] = c.x; K.mu[4 * g + 1] = c.y; K.mu[4 * g + 2] = c.z; K.mu[4 * g + 3] = c.w;
      }
    };
    for (int s = 0; s <= my_imgs; ++s) {
      const bool do1 = s < my_imgs, do2 = s >= 1;
      const int buf1 = s & 1;
      if (do1) mbar_wait(band_empty + buf1, ((s >> 1) & 1) ^ 1);  // conv2 has finished with this band's previous image
      const int64_t img2 = (int64_t)blockIdx.x + (int64_t)(s - 1) * gridDim.x;
#pragma unroll 1
      for (int t = set; t < TILES; t += 2) {
        const int r0 = (t >> 2) * 16 + 4 * quarter, c = (t & 3) * 8 + j;
        if (do1) {
          // ---- conv1 block of image s: requantise, store into the conv2 band (position (r+1)*33 + (c+1), 64-byte rows,
          //      16-byte chunk q4 ^ swizzle(position))
          load_consts(0);
          const int acc_it = s * TILES + t;
          const uint32_t slot = acc_it % SLOTS;
          const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * COUT;
          uint8_t* band = a2_smem + buf1 * A2_BYTES;
          auto store = [&](int half, int ss, const uint32_t (&packed)[4]) {
            const int pos = (r0 + 2 * half + ss + 1) * P2 + c + 1;
            *reinterpret_cast<uint4*>(band + pos * CIN2 + ((q4 ^ ((pos >> 1) & 3)) << 4)) =
                make_uint4(packed[0], packed[1], packed[2], packed[3]);
          };
          auto release = [&]() {
            if (lane == 0) mbar_arrive(t1_empty + slot);
          };
          mbar_wait(t1_full + slot, (acc_it / SLOTS) & 1);
          tc_fence_after();
          epi_block_store16<CHECK1>(t_addr, K, L1, ch0, fast1, args.zp1, args.lo1, store, release);
        }
        if (do2) {
          // ---- conv2 block of image s-1: 2x2 max-pool, requantise, store to global
          load_consts(1);
          const int acc_it = (s - 1) * TILES + t;
          const uint32_t slot = acc_it % SLOTS;
          const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (SLOTS + slot) * COUT;
          uint8_t* out = args.y + ((img2 * (IMG / 2) + (r0 >> 1) + (j & 1)) * (IMG / 2) + (c >> 1)) * (int64_t)COUT + ch0;
          auto release = [&]() {
            if (lane == 0) mbar_arrive(t2_empty + slot);
          };
          mbar_wait(t2_full + slot, (acc_it / SLOTS) & 1);
          tc_fence_after();
          epi_block_pool16_units<CHECK2>(t_addr, K, L2, ch0, fast2, args.zp2, args.lo2, out, true, lane, release);
        }
      }
      if (do1) {  // this warp's part of image s is in the band
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(band_full + buf1);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PROD_WARP0) {
    __syncwarp();
    tmem_dealloc(tmem_base, 2 * f12::SLOTS * f12::COUT);
  }
}

}  // namespace b200q

using namespace b200q;

// Fused quantize + conv1 + ReLU + conv2 + ReLU + 2x2 max-pool.  Returns B200Q_ERR_INVALID_ARG for geometries / constants
// the kernel does not cover (callers fall back to the two separate entry points).
extern "C" int b200q_conv12_fused(const float* x, uint8_t* y, int64_t b, float inv_scale, const b200q_conv3x3* L1,
                                  const b200q_conv3x3* L2, void* stream) {
  B200Q_REQUIRE(L1 && L2 && ((x && y) || b == 0), "conv12_fused: null pointer");
  B200Q_REQUIRE(L1->cin == 4 && L1->cout == 64 && L1->img == 32 && L2->cin == 64 && L2->cout == 64 && L2->img == 32,
                "conv12_fused: unsupported geometry");
  B200Q_REQUIRE(L1->w && L2->w && L1->corr_host && L2->corr_host && L1->rq.mult_host && L1->rq.bdiv_host &&
                    L2->rq.mult_host && L2->rq.bdiv_host,
                "conv12_fused: layers must carry host mirrors of their constants");
  B200Q_REQUIRE((L1->rq.flags & B200Q_RQ_BOUNDED) != 0, "conv12_fused: conv1 constants must be B200Q_RQ_BOUNDED");
  B200Q_REQUIRE(L2->zp_x == L1->rq.zp_out, "conv12_fused: conv2 input zero-point must equal conv1 output zero-point");
  B200Q_REQUIRE((uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0 && (uintptr_t)L2->w % 16 == 0,
                "conv12_fused: buffers must be 16-byte aligned");
  if (b == 0) return 0;
  F12Consts consts;
  const b200q_conv3x3* Ls[2] = {L1, L2};
  for (int l = 0; l < 2; ++l)
    for (int c = 0; c < f12::COUT; ++c) {
      const int32_t corr = Ls[l]->corr_host[4 * f12::COUT + c];  // class 4 = all nine taps (pads hold the zero-point)
      consts.cm[l][c] = (int32_t)(MAGIC_BITS - (uint32_t)corr);
      consts.k1[l][c] = -(MAGIC_F + (float)corr);
      consts.bdiv[l][c] = Ls[l]->rq.bdiv_host[c];
      consts.mult[l][c] = Ls[l]->rq.mult_host[c];
    }
  const bool check1 = !(L1->rq.flags & B200Q_RQ_ACC22);
  const bool check2 = !((L2->rq.flags & B200Q_RQ_BOUNDED) && (L2->rq.flags & B200Q_RQ_ACC22));
  F12Args args{x,
               y,
               L1->w,
               L2->w,
               b,
               inv_scale,
               L1->zp_x,
               L1->rq.zp_out,
               L1->rq.relu ? L1->rq.zp_out : 0,
               L2->rq.zp_out,
               L2->rq.relu ? L2->rq.zp_out : 0,
               1,
               (L2->rq.flags & B200Q_RQ_BOUNDED) ? 1 : 0};
  void (*kernel)(F12Consts, F12Args) =
      check1 ? (check2 ? conv12_fused_kernel<true, true> : conv12_fused_kernel<true, false>)
             : (check2 ? conv12_fused_kernel<false, true> : conv12_fused_kernel<false, false>);
  static uint64_t attr_mask[4] = {0, 0, 0, 0};
  const int idx = (check1 ? 2 : 0) + (check2 ? 1 : 0);
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), f12::SMEM, &attr_mask[idx])) return rc;
  const int grid = b < num_sms() ? (int)b : num_sms();
  kernel<<<grid, f12::THREADS, f12::SMEM, (cudaStream_t)stream>>>(consts, args);
  return launched("conv12_fused_kernel");
}
