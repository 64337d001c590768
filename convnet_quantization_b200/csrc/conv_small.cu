// 3x3 convolution for a handful of images (the whole-network executor at batch 1..CONV_TINY_MAX_B): CUDA cores, dp4a.
//
// At batch 1 a layer is ~38 M MACs - nothing - and its duration is the length of its chain of dependent round trips.
// The tensor-core small-batch shape (igemm_tc.cu, N tile 64) still gives each of 4..16 CTAs its whole K to stream:
// 9..18 TMA chunks through an 8-deep ring, 18..72 dependent MMAs, a TMEM round trip (4.1..7.6 us per layer, r02 graph
// breakdown).  Here the same layer is 64 (conv2: 128) CTAs per image that each do ONE round trip:
//   * a CTA owns 8 image rows (4 at 32x32) x 4 output channels.  Its four weight rows live in REGISTERS, K split across the 32 lanes
//     of every warp (lane l holds words l, l+32, ... of each row) and are loaded BEFORE the programmatic-dependent-launch
//     wait, i.e. while the previous layer still runs;
//   * after the wait the (rows+2) x (IMG+2) pixel halo tile is copied to shared memory in one batch of independent 16-byte
//     loads per thread; border pixels hold the input zero-point (real-domain zero), so one correction constant per
//     channel (the interior class of b200q_conv3x3.corr) is exact for every pixel;
//   * a warp walks its pixels: per pixel W conflict-free shared loads and 4 W dp4a per lane, four REDUX warp sums; every
//     eight pixels the 32 (pixel, channel) sums sit one per lane and are requantised with the exact fbgemm form
//     (requant_u8) in one go; bytes are staged in shared memory, max-pooled there when the layer pools, and leave as
//     4-byte stores.
// Integer accumulation is exact in any order and the requantisation is the reference form, so the result is
// bit-identical to the tensor-core kernels (tests/test_gpu_conv.py::test_conv_tiny_batches).
#include "common.cuh"

namespace b200q {
namespace {

constexpr int CS_THREADS = 256;
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int CS_CO = 4;    // output channels per CTA

template <int IMG, int CIN>
struct CsCfg {
  static constexpr int ROWS = IMG == 32 ? 4 : 8;     // image rows per CTA (32x32: 128 CTAs per image, 16 pixels per warp)
  static constexpr int CINW = CIN / 4;               // 32-bit words per pixel
  static constexpr int KW = 9 * CINW;                // words per weight row
  static constexpr int W = (KW + 31) / 32;           // words per lane and row
  static constexpr int PW = IMG + 2, PH = ROWS + 2;
  static constexpr int TILE_WORDS = PH * PW * CINW;
  static constexpr int CH16 = CIN / 16;              // 16-byte chunks per pixel
  static constexpr int TILE_CHUNKS = PH * PW * CH16;
  static constexpr int LOADS = (TILE_CHUNKS + CS_THREADS - 1) / CS_THREADS;
  static constexpr int PIX = ROWS * IMG;          // output pixels per CTA before pooling
  static constexpr int PIX_PER_WARP = PIX / CS_WARPS;
  static constexpr int SLICES = IMG / ROWS;
  static constexpr int UNROLL = W <= 9 ? 8 : 4;      // independent pixels in flight per warp (register budget)
  static_assert(CIN % 16 == 0 && IMG % ROWS == 0 && ROWS % 2 == 0 && PIX_PER_WARP % 8 == 0, "geometry");
  static_assert(TILE_WORDS * 4 + PIX * CS_CO <= 48 * 1024, "static shared memory");
};

struct CsArgs {
  const uint8_t* x;
  uint8_t* y;
  const int8_t* w;
  const int32_t* corr;  // interior border class: zp_x * sum of all taps
  const float* mult;
  const float* bdiv;
  int zp_x, zp_out, lo;
};

template <int IMG, int CIN, int COUT, bool POOL>
__global__ void __launch_bounds__(CS_THREADS) conv3x3_tiny_kernel(const CsArgs a) {
  using C = CsCfg<IMG, CIN>;
  __shared__ __align__(16) uint32_t s_x[C::TILE_WORDS];
  __shared__ __align__(16) uint8_t s_out[C::PIX * CS_CO];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int COGS = COUT / CS_CO;
  const int cog = blockIdx.x % COGS;
  const int slice = (blockIdx.x / COGS) % C::SLICES;
  const int img = blockIdx.x / (COGS * C::SLICES);
  const int co0 = cog * CS_CO, r0 = slice * C::ROWS;
  pdl_launch_dependents();

  // ---- independent of the previous layer: this CTA's weights (K split across the lanes) and channel constants
  uint32_t wr[CS_CO][C::W];
  {
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(a.w) + (int64_t)co0 * C::KW;
#pragma unroll
    for (int c = 0; c < CS_CO; ++c)
#pragma unroll
      for (int s = 0; s < C::W; ++s) {
        const int j = lane + 32 * s;
        wr[c][s] = j < C::KW ? __ldg(w32 + c * C::KW + j) : 0u;
      }
  }
  int off[C::W];  // word offset of lane's K word s inside the halo tile, relative to the output pixel's tap (0,0)
#pragma unroll
  for (int s = 0; s < C::W; ++s) {
    const int j = lane + 32 * s;
    const int tap = j / C::CINW, cw = j % C::CINW;
    off[s] = tap < 9 ? ((tap / 3) * C::PW + tap % 3) * C::CINW + cw : cw;  // words past K: weight 0, any valid address
  }
  const int my_c = co0 + (lane & 3);
  const int corr = __ldg(a.corr + my_c);
  const float mult = __ldg(a.mult + my_c), bdiv = __ldg(a.bdiv + my_c);

  pdl_wait();  // x is the previous kernel's output; y may still be read by it

  // ---- halo tile -> shared memory: all of a thread's loads in flight at once
  {
    const uint8_t* ximg = a.x + (int64_t)img * IMG * IMG * CIN;
    const uint32_t zp1 = (uint32_t)a.zp_x * 0x01010101u;
    uint4 v[C::LOADS];
#pragma unroll
    for (int k = 0; k < C::LOADS; ++k) {
      const int i = tid + k * CS_THREADS;
      const int pix = i / C::CH16, ch = i % C::CH16;
      const int gy = r0 - 1 + pix / C::PW, gx = pix % C::PW - 1;
      v[k] = make_uint4(zp1, zp1, zp1, zp1);
      if (i < C::TILE_CHUNKS && gy >= 0 && gy < IMG && gx >= 0 && gx < IMG)
        v[k] = __ldg(reinterpret_cast<const uint4*>(ximg + (int64_t)(gy * IMG + gx) * CIN) + ch);
    }
#pragma unroll
    for (int k = 0; k < C::LOADS; ++k) {
      const int i = tid + k * CS_THREADS;
      if (i < C::TILE_CHUNKS) reinterpret_cast<uint4*>(s_x)[i] = v[k];
    }
  }
  __syncthreads();

  // ---- the warp's pixels, eight at a time: 32 (pixel, channel) sums, one per lane
#pragma unroll 1
  for (int g = 0; g < C::PIX_PER_WARP / 8; ++g) {
    int keep = 0;
#pragma unroll C::UNROLL
    for (int i = 0; i < 8; ++i) {
      const int p = warp * C::PIX_PER_WARP + g * 8 + i;
      const uint32_t* base = s_x + ((p / IMG) * C::PW + p % IMG) * C::CINW;
      int acc[CS_CO] = {};
#pragma unroll
      for (int s = 0; s < C::W; ++s) {
        const uint32_t xv = base[off[s]];
#pragma unroll
        for (int c = 0; c < CS_CO; ++c) acc[c] = dp4a_us(xv, wr[c][s], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < CS_CO; ++c) {
        const int t = __reduce_add_sync(0xffffffffu, acc[c]);
        if (lane == i * CS_CO + c) keep = t;
      }
    }
    const int p = warp * C::PIX_PER_WARP + g * 8 + (lane >> 2);
    s_out[p * CS_CO + (lane & 3)] = (uint8_t)requant_u8(keep - corr, bdiv, mult, a.zp_out, a.lo);
  }
  __syncthreads();

  // ---- write out (4 channels = one word per pixel), pooling 2x2 windows first when the layer pools
  const uint32_t* so = reinterpret_cast<const uint32_t*>(s_out);
  if constexpr (POOL) {
    constexpr int O = IMG / 2;
    for (int t = tid; t < C::PIX / 4; t += CS_THREADS) {
      const int qy = t / O, qx = t % O;
      const int p00 = (2 * qy) * IMG + 2 * qx;
      const uint32_t m = max4_u8x4(so[p00], so[p00 + 1], so[p00 + IMG], so[p00 + IMG + 1]);
      uint8_t* dst = a.y + ((int64_t)(img * O + r0 / 2 + qy) * O + qx) * COUT + co0;
      *reinterpret_cast<uint32_t*>(dst) = m;
    }
  } else {
    for (int t = tid; t < C::PIX; t += CS_THREADS) {
      uint8_t* dst = a.y + ((int64_t)(img * IMG + r0 + t / IMG) * IMG + t % IMG) * COUT + co0;
      *reinterpret_cast<uint32_t*>(dst) = so[t];
    }
  }
}

template <int IMG, int CIN, int COUT>
int launch_tiny(const uint8_t* x, uint8_t* y, int64_t b, const b200q_conv3x3* L, bool pool, cudaStream_t s) {
  using C = CsCfg<IMG, CIN>;
  const CsArgs a{x, y, L->w, L->corr + 4 * COUT, L->rq.mult, L->rq.bdiv, L->zp_x, L->rq.zp_out,
                 L->rq.relu ? L->rq.zp_out : 0};
  const int grid = (int)b * C::SLICES * (COUT / CS_CO);
  if (pool) return launch_kernel("conv3x3_tiny_kernel", conv3x3_tiny_kernel<IMG, CIN, COUT, true>, grid, CS_THREADS, 0, s, a);
  return launch_kernel("conv3x3_tiny_kernel", conv3x3_tiny_kernel<IMG, CIN, COUT, false>, grid, CS_THREADS, 0, s, a);
}

}  // namespace

int conv3x3_tiny_dispatch(const uint8_t* x, uint8_t* y, int64_t b, const b200q_conv3x3* L, bool pool, cudaStream_t s,
                          int* rc) {
#define B200Q_TINY_CASE(IMG_, CIN_, COUT_)                            \
  if (L->img == IMG_ && L->cin == CIN_ && L->cout == COUT_) {         \
    *rc = launch_tiny<IMG_, CIN_, COUT_>(x, y, b, L, pool, s);        \
    return 0;                                                         \
  }
  B200Q_TINY_CASE(32, 64, 64)
  B200Q_TINY_CASE(16, 64, 128)
  B200Q_TINY_CASE(16, 128, 128)
  B200Q_TINY_CASE(8, 128, 256)
  B200Q_TINY_CASE(8, 256, 256)
#undef B200Q_TINY_CASE
  return 1;
}

}  // namespace b200q
