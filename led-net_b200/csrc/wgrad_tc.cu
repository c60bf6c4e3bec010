// Weight gradient of a 3x3 or 1x1 convolution as an implicit GEMM on the 5th-gen tensor cores (north_star kernel 6).
//
// Replaces what the reference gets from cuDNN's convolution_backward through autograd
//   (mmseg/models/segmentors/encoder_decoder.py:161-185 loss -> decode_heads/led_head.py:101-146), for every
//   Conv2d(3x3 padding 1 | 1x1, stride 1 | 2) of the trunk and head whose channel counts are multiples of 32.
// One kind::tf32 pass by default (a weight gradient is a leaf of the backward pass; see train.cu g_wgrad_passes), three
// error-compensated passes (X3, 3x3 only) on request.
//
//   dW[co][ci][kh][kw] = sum over (n, oh, ow) of dY[n, oh, ow, co] * X[n, oh*s + kh - 1, ow*s + kw - 1, ci]
//
// GEMM view: the REDUCTION dimension is the pixel, so both operands are read "MN-major": a shared-memory row is one pixel,
// its 128 bytes are 32 fp32 channels - exactly what TMA writes for an NHWC box of 32 channels (128 B swizzle, 32 B atoms).  One
// tcgen05.mma kind::tf32 consumes K = 8 rows = the 8 pixels of one tile row:
//   B (N side)  dY tile rows [r*8, r*8+8) of the 16 x 8 output tile, N = 32..128 output channels (32-channel slabs LBO apart)
//   A (M side)  the X halo slab of one 32-channel block, window starting at pixel (r*s' + kh, 0): M = 128 = FOUR 32-channel
//               atoms whose stride (LBO) is ONE PIXEL ROW (128 B) - atom a is the same window shifted by a pixels, i.e. filter
//               column kw = a.  Rows 0..95 of the accumulator are (kw, ci) for kw = 0, 1, 2; rows 96..127 (a fourth shift)
//               are never read.  The swizzle XOR is applied on absolute shared-memory address bits (probed for conv_tc.cu),
//               so overlapping, unaligned atoms read exactly what TMA wrote.
//   D           fp32 in TMEM: one accumulator [128 x N] per (input-channel block, kh); all accumulators of a CTA live in TMEM
//               for its whole life (<= 512 columns) and are drained ONCE at the end.
// Stride 2: X is viewed as [N][H/2][2][W/2][2C] (row / column parity split, as conv_tc.cu does); a filter row kh picks a row
// parity, filter columns 0 and 2 are two shifts of the odd-column slab (atoms 0, 1), column 1 the even-column slab (atom 0),
// each pair with its own accumulator.
//
// Work split: a CTA owns (input-channel blocks, output-channel blocks) = a "unit" and every G-th pixel tile of it; it
// writes its accumulators to its own slot of the workspace and wgrad_tc_sum_kernel adds the G slots of a unit in index
// order into the OIHW gradient: no atomics, bit-reproducible.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-5 drain TMEM at the end.
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "tc_common.cuh"

namespace ledb {
namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int TH = 16, TW = 8;
constexpr uint32_t SMEM_MAX = 227 * 1024;

struct WgParams {
  int s2;                  // stride 2
  int ncc, nco;            // 32-channel blocks of X / dY per CTA
  int ncols;               // nco * 32 = N of every MMA
  int nacc;                // accumulators per CTA: ncc * 3 (stride 1) or ncc * 6 (stride 2)
  int units_ci, units_co;  // number of units along Cin / Cout
  int G;                   // CTAs per unit
  int tiles_w, tiles_h, N;
  int64_t tiles;           // pixel tiles (per unit)
  int Cin, Cout;
  int slab_w;              // X slab width in pixels (10 stride 1, 9 stride 2)
  uint32_t xslab_bytes, yslab_bytes, stage_bytes, tx_bytes;
  int nstages;
  uint32_t tmem_cols;
  int in_ld2;              // stride 2: pixel stride of X in 2-byte units (column-parity offset of the 5-D view)
  int k1;                  // 1x1 filter: no halo, the four M atoms are four consecutive 32-channel blocks of X (LBO = one slab)
  int prow;                // accumulator rows a CTA writes per accumulator: 96 (3x3: three filter columns) or 128 (1x1)
  float* part;             // [units * G][nacc][96][ncols]
};

// X3: error-compensated three-pass mode (see conv_tc.cu): both operands are activations here, so warps 6..9 write the low
// parts x - trunc_tf32(x) of EVERY slab of a stage (X and dY) into the stage's second half, and each product becomes
//   X_raw * dY_raw + X_lo * dY_raw + X_raw * dY_lo      (the tensor core truncates the raw operands itself)
// K1: 1x1 filters (single-pass mode only).  X tile = the output tile's own pixels (stride 2: the even-even parity slab), no
// taps; the four 32-channel atoms of M are four consecutive input-channel blocks, one slab apart, so one MMA covers 128 input
// channels x N output channels.  Stages always reserve four X slabs: atoms past the last real block read the next slab's
// bytes (finite or not, those accumulator rows are never read).
template <bool S2, bool X3, bool K1 = false>
__global__ void __launch_bounds__(X3 ? kThreads + 128 : kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                const __grid_constant__ WgParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P.nstages * P.stage_bytes);
  uint64_t* full = bars;          // [8]
  uint64_t* empty = bars + 8;     // [8]
  uint64_t* done = bars + 16;     // [1]
  uint64_t* lo_full = bars + 17;  // [8] X3: low parts of stage i written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x / P.G, g = blockIdx.x % P.G;
  const int uci = unit % P.units_ci, uco = unit / P.units_ci;
  const int nx = K1 ? 1 : (S2 ? 4 : 1);      // X slabs per channel block (stride 2: row parity x column parity)
  const int xregion = K1 ? 4 : P.ncc * nx;   // X slabs a stage reserves

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmY);
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&lo_full[i], 4); }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0, ph = 0;
      for (int64_t t = g; t < P.tiles; t += P.G) {
        const int tw = (int)(t % P.tiles_w), th = (int)((t / P.tiles_w) % P.tiles_h), n = (int)(t / ((int64_t)P.tiles_w * P.tiles_h));
        const int h0 = th * TH, w0 = tw * TW;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], P.tx_bytes);
        uint8_t* st = stages + (size_t)s * P.stage_bytes;
        const uint32_t bar = smem_u32(&full[s]);
        for (int c = 0; c < P.ncc; ++c) {
          const int ch = (uci * P.ncc + c) * 64;                  // 2-byte units: 32 fp32 channels = 64 units
          if (K1) {
            if (S2) tma_load_5d(smem_u32(st + (size_t)c * P.xslab_bytes), &tmX, bar, ch, w0, 0, h0, n);
            else tma_load_4d(smem_u32(st + (size_t)c * P.xslab_bytes), &tmX, bar, ch, w0, h0, n);
          } else if (S2) {
            // slab (pr, pc): rows of parity pr, columns of parity pc.  Odd rows / columns start one half-pixel earlier
            // (filter tap 0 reads input 2*o - 1).
            for (int q = 0; q < 4; ++q) {
              const int pr = q >> 1, pc = q & 1;
              tma_load_5d(smem_u32(st + (size_t)(c * 4 + q) * P.xslab_bytes), &tmX, bar, pc * P.in_ld2 + ch,
                          w0 - pc, pr, h0 - pr, n);
            }
          } else {
            tma_load_4d(smem_u32(st + (size_t)c * P.xslab_bytes), &tmX, bar, ch, w0 - 1, h0 - 1, n);
          }
        }
        uint8_t* sy = st + (size_t)xregion * P.xslab_bytes;
        for (int o = 0; o < P.nco; ++o)
          tma_load_4d(smem_u32(sy + (size_t)o * P.yslab_bytes), &tmY, bar, (uco * P.nco + o) * 64, w0, h0, n);
        if (++s == P.nstages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32_m128(P.ncols, 1u, 1u);
      // MN-major tf32 operands exist in ONE shared-memory layout (cute: Layout_MN_SW128_32B_Atom, descriptor layout type 1,
      // TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 128 B rows (32 channels) whose 32-byte chunks are XORed with the row
      // index mod 4; a K atom is 4 rows, so K = 8 spans two atoms SBO = 512 B apart.  LBO = stride between 32-channel
      // atoms - A: one pixel row (filter columns); B: one dY slab.
      const uint32_t a_hi = desc_hi(512, 1u), b_hi = desc_hi(512, 1u);
      const uint32_t a_lbo = (128u >> 4) << 16, b_lbo = (P.yslab_bytes >> 4) << 16;
      int s = 0, ph = 0;
      uint32_t first = 0;                                          // 0 until the accumulators hold a first product
      const uint32_t half = X3 ? P.stage_bytes / 2 : 0u;           // offset of the low-part copy of a stage
      for (int64_t t = g; t < P.tiles; t += P.G) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t st = smem_u32(stages + (size_t)s * P.stage_bytes);
        const uint32_t sy = st + (uint32_t)xregion * P.xslab_bytes;
        // pass 0: raw x raw; X3 passes 1, 2 (after the converter warps have arrived): X_lo x dY_raw, X_raw x dY_lo
#pragma unroll 1
        for (int pass = 0; pass < (X3 ? 3 : 1); ++pass) {
          if (X3 && pass == 1) { mbar_wait(&lo_full[s], ph); tc_fence_after(); }
          const uint32_t xo = (pass == 1) ? half : 0u, yo = (pass == 2) ? half : 0u;
          for (int r = 0; r < TH; ++r) {
            const uint32_t b_lo = (((sy + yo + (uint32_t)(r * TW) * 128u) >> 4) & 0x3FFFu) | b_lbo;
            // passes 1, 2 accumulate into their own columns (lo_cols after the main ones): pass 1 opens them on the first tile
            const uint32_t acc = pass == 2 ? 1u : (first | (uint32_t)r);
            const uint32_t dcol = pass ? (uint32_t)(P.nacc * P.ncols) : 0u;
            if (K1) {
              const uint32_t a_lo = (((st + (uint32_t)(r * TW) * 128u) >> 4) & 0x3FFFu) | ((P.xslab_bytes >> 4) << 16);
              tc_mma2_tf32(tmem_base, a_lo, a_hi, b_lo, b_hi, idesc, acc);
            } else
            for (int c = 0; c < P.ncc; ++c) {
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
                if (S2) {
                  // input row 2*oh + kh - 1: kh = 0 -> odd rows, slab row r (slab starts at half-row oh0 - 1);
                  // kh = 1 -> even rows, slab row r; kh = 2 -> odd rows, slab row r + 1
                  const int pr = (kh == 1) ? 0 : 1, srow = (kh == 2) ? r + 1 : r;
#pragma unroll
                  for (int pc = 0; pc < 2; ++pc) {
                    // pc = 1: atoms 0, 1 = filter columns 0, 2 (odd columns ow - 1, ow); pc = 0: atom 0 = filter column 1
                    const uint32_t xs = st + xo + (uint32_t)(c * 4 + pr * 2 + pc) * P.xslab_bytes;
                    const uint32_t a_lo = (((xs + (uint32_t)(srow * P.slab_w) * 128u) >> 4) & 0x3FFFu) | a_lbo;
                    const uint32_t d = tmem_base + dcol + (uint32_t)(((c * 3 + kh) * 2 + pc) * P.ncols);
                    tc_mma2_tf32(d, a_lo, a_hi, b_lo, b_hi, idesc, acc);
                  }
                } else {
                  const uint32_t xs = st + xo + (uint32_t)c * P.xslab_bytes;
                  const uint32_t a_lo = (((xs + (uint32_t)((r + kh) * P.slab_w) * 128u) >> 4) & 0x3FFFu) | a_lbo;
                  const uint32_t d = tmem_base + dcol + (uint32_t)((c * 3 + kh) * P.ncols);
                  tc_mma2_tf32(d, a_lo, a_hi, b_lo, b_hi, idesc, acc);
                }
              }
            }
          }
        }
        tc_commit(&empty[s]);
        first = 1;
        if (++s == P.nstages) { s = 0; ph ^= 1; }
      }
      tc_commit(done);
    }
    __syncwarp();
  } else if (warp >= 6) {
    if constexpr (X3) {
      // low-part converter: element-wise over the raw bytes of the stage's first half (the swizzle only permutes 16-byte chunks)
      const uint32_t ctid = threadIdx.x - 6 * 32, half = P.stage_bytes / 2;
      int s = 0, ph = 0;
      for (int64_t t = g; t < P.tiles; t += P.G) {
        mbar_wait(&full[s], ph);
        const uint32_t src = smem_u32(stages + (size_t)s * P.stage_bytes), dst = src + half;
        for (uint32_t o = ctid * 16; o < half; o += 128 * 16) {
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src + o) : "memory");
          const uint32_t a = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & 0xFFFFE000u));
          const uint32_t b = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & 0xFFFFE000u));
          const uint32_t c = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & 0xFFFFE000u));
          const uint32_t d = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & 0xFFFFE000u));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + o), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&lo_full[s]);
        if (++s == P.nstages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ---- drain: warp q owns TMEM lanes [32q, 32q + 32) = accumulator rows of filter column q (stride 1); the fourth
    // quadrant holds the unused shift.  part[cta][acc][row 0..95][ncols]
    const int q = warp & 3;
    const bool has_work = g < P.tiles;                 // a CTA without tiles never touched its accumulators: write zeros
    if (has_work) { mbar_wait(done, 0); tc_fence_after(); }
    if (q * 32 < P.prow) {
      float* mine = P.part + (size_t)blockIdx.x * P.nacc * P.prow * P.ncols;
      for (int a = 0; a < P.nacc; ++a) {
        float* row = mine + ((size_t)a * P.prow + q * 32 + lane) * P.ncols;
        for (int c0 = 0; c0 < P.ncols; c0 += 16) {
          uint32_t v[16];
          if (has_work) {
            tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * P.ncols + c0), v);
            tc_wait_ld();
            if (X3) {   // main + correction accumulator
              uint32_t u[16];
              tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((P.nacc + a) * P.ncols + c0), u);
              tc_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<uint4*>(row + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, P.tmem_cols);
  }
}

// dW[co][ci][kh][kw] = sum over the G slots of the unit, in slot order
__global__ void __launch_bounds__(256)
wgrad_tc_sum_kernel(const float* __restrict__ part, float* __restrict__ dw, int Cin, int Cout, int ncc, int nco, int units_ci,
                    int G, int s2) {
  const int ncols = nco * 32;
  const int nacc = ncc * (s2 ? 6 : 3);
  const int64_t per_cta = (int64_t)nacc * 96 * ncols;
  const int64_t total = (int64_t)Cout * Cin * 9;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    // walk the OUTPUT in an order whose fastest index is the output channel (the partials' contiguous axis)
    const int co = (int)(i % Cout);
    const int64_t r = i / Cout;
    const int kw = (int)(r % 3), kh = (int)((r / 3) % 3), ci = (int)(r / 9);
    const int uco = co / ncols, col = co % ncols;
    const int cb = ci / 32, uci = cb / ncc, c = cb % ncc, cil = ci % 32;
    int acc, row;
    if (s2) { const int pc = (kw == 1) ? 0 : 1; acc = (c * 3 + kh) * 2 + pc; row = (kw == 2 ? 32 : 0) + cil; }
    else { acc = c * 3 + kh; row = kw * 32 + cil; }
    const int unit = uco * units_ci + uci;
    const float* p = part + (int64_t)unit * G * per_cta + ((int64_t)acc * 96 + row) * ncols + col;
    // four independent partial sums (slots k, k+1, k+2, k+3 in turn): four loads in flight instead of a dependent chain of
    // up to 148; the grouping depends on G only, so the result is as reproducible as the plain loop's
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 4 <= G; k += 4) {
      s0 += p[(int64_t)k * per_cta]; s1 += p[(int64_t)(k + 1) * per_cta];
      s2 += p[(int64_t)(k + 2) * per_cta]; s3 += p[(int64_t)(k + 3) * per_cta];
    }
    for (; k < G; ++k) s0 += p[(int64_t)k * per_cta];
    const float s = (s0 + s1) + (s2 + s3);
    dw[(((int64_t)co * Cin + ci) * 3 + kh) * 3 + kw] = s;
  }
}

// 1x1: dW[co][ci] = sum over the G slots of the unit; accumulator row = (block inside the unit's group) * 32 + ci % 32
__global__ void __launch_bounds__(256)
wgrad_tc_sum1_kernel(const float* __restrict__ part, float* __restrict__ dw, int Cin, int Cout, int ncc, int nco, int units_ci,
                     int G) {
  const int ncols = nco * 32;
  const int64_t per_cta = (int64_t)128 * ncols;
  const int64_t total = (int64_t)Cout * Cin;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int co = (int)(i % Cout), ci = (int)(i / Cout);
    const int uco = co / ncols, col = co % ncols;
    const int cb = ci / 32, uci = cb / ncc, c = cb % ncc;
    const int unit = uco * units_ci + uci;
    const float* p = part + (int64_t)unit * G * per_cta + (int64_t)(c * 32 + ci % 32) * ncols + col;
    // four independent partial sums (slots k, k+1, k+2, k+3 in turn): four loads in flight instead of a dependent chain of
    // up to 148; the grouping depends on G only, so the result is as reproducible as the plain loop's
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 4 <= G; k += 4) {
      s0 += p[(int64_t)k * per_cta]; s1 += p[(int64_t)(k + 1) * per_cta];
      s2 += p[(int64_t)(k + 2) * per_cta]; s3 += p[(int64_t)(k + 3) * per_cta];
    }
    for (; k < G; ++k) s0 += p[(int64_t)k * per_cta];
    const float s = (s0 + s1) + (s2 + s3);
    dw[(int64_t)co * Cin + ci] = s;
  }
}

struct WgPlan {
  WgParams P;
  int grid;
  size_t smem;
  int64_t part_floats;
  bool ok, x3;
};

int num_sms() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

WgPlan plan(int N, int H, int W, int Cin, int Cout, int k, int stride, int passes) {
  const bool x3 = passes == 3;
  WgPlan pl{};
  pl.ok = false;
  if ((k != 3 && k != 1) || (stride != 1 && stride != 2)) return pl;
  if (Cin < 32 || Cin % 32 || Cout < 32 || Cout % 32) return pl;
  if (stride == 2 && ((H & 1) || (W & 1))) return pl;
  const int pad = k / 2;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  if (Ho % TH || Wo % TW) return pl;
  WgParams& P = pl.P;
  P.s2 = stride == 2;
  const int bi = Cin / 32, bo = Cout / 32;
  if (k == 1) {
    if (x3) return pl;                                   // single-pass mode only
    if (bi > 4 && bi % 4) return pl;
    P.k1 = 1; P.prow = 128;
    P.ncc = std::min(bi, 4);
    P.nco = std::min(bo, 2);
    while (bo % P.nco) --P.nco;
    P.ncols = P.nco * 32;
    P.nacc = 1;
    P.units_ci = bi / P.ncc; P.units_co = bo / P.nco;
    const int units = P.units_ci * P.units_co;
    P.tiles_w = Wo / TW; P.tiles_h = Ho / TH; P.N = N;
    P.tiles = (int64_t)N * P.tiles_w * P.tiles_h;
    P.G = (int)std::max<int64_t>(1, std::min<int64_t>(P.tiles, num_sms() / units));
    P.Cin = Cin; P.Cout = Cout;
    P.slab_w = TW;
    P.xslab_bytes = TH * TW * 128;
    P.yslab_bytes = TH * TW * 128;
    P.stage_bytes = 4u * P.xslab_bytes + (uint32_t)P.nco * P.yslab_bytes;
    P.tx_bytes = (uint32_t)P.ncc * P.xslab_bytes + (uint32_t)P.nco * P.yslab_bytes;
    P.nstages = (int)std::min<uint32_t>(6, (SMEM_MAX - 1024 - 256) / P.stage_bytes);
    if (P.nstages < 2) return pl;
    uint32_t cols = 32;
    while (cols < (uint32_t)P.ncols) cols <<= 1;
    P.tmem_cols = cols;
    P.in_ld2 = Cin * 2;
    pl.grid = units * P.G;
    pl.smem = 1024 + (size_t)P.nstages * P.stage_bytes + 256;
    pl.x3 = false;
    pl.part_floats = (int64_t)pl.grid * 128 * P.ncols;
    pl.ok = true;
    return pl;
  }
  P.k1 = 0; P.prow = 96;
  const int acc_per_c = P.s2 ? 6 : 3;
  // output-channel blocks per CTA: up to 4 (N = 128); stride 2 keeps 2 so that two stages of four X slabs fit
  P.nco = std::min(bo, (P.s2 && x3) ? 1 : ((P.s2 || x3) ? 2 : 4));   // three-pass mode: every slab and accumulator exists twice
  while (bo % P.nco) --P.nco;
  P.ncols = P.nco * 32;
  P.ncc = std::max(1, std::min(bi, 512 / (acc_per_c * P.ncols)));
  P.ncc = std::min(P.ncc, (P.s2 || x3) ? 1 : 2);
  while (bi % P.ncc) --P.ncc;
  P.nacc = P.ncc * acc_per_c;
  P.units_ci = bi / P.ncc; P.units_co = bo / P.nco;
  const int units = P.units_ci * P.units_co;
  P.tiles_w = Wo / TW; P.tiles_h = Ho / TH; P.N = N;
  P.tiles = (int64_t)N * P.tiles_w * P.tiles_h;
  P.G = (int)std::max<int64_t>(1, std::min<int64_t>(P.tiles, num_sms() / units));
  P.Cin = Cin; P.Cout = Cout;
  P.slab_w = P.s2 ? TW + 1 : TW + 2;
  const int slab_h = P.s2 ? TH + 1 : TH + 2;
  const uint32_t xbox = (uint32_t)(slab_h * P.slab_w * 128);
  P.xslab_bytes = (xbox + 512 + 1023) / 1024 * 1024;      // + the fourth (unused) atom's over-read of up to three pixel rows
  P.yslab_bytes = TH * TW * 128;
  const int nx = P.s2 ? 4 : 1;
  P.stage_bytes = ((uint32_t)(P.ncc * nx) * P.xslab_bytes + (uint32_t)P.nco * P.yslab_bytes) * (x3 ? 2u : 1u);
  P.tx_bytes = (uint32_t)(P.ncc * nx) * xbox + (uint32_t)P.nco * P.yslab_bytes;
  const uint32_t avail = SMEM_MAX - 1024 - 256;
  P.nstages = (int)std::min<uint32_t>(6, avail / P.stage_bytes);
  if (P.nstages < (x3 ? 1 : 2)) return pl;       // stride 2 in three-pass mode runs a single stage (load, convert, multiply in turn)
  uint32_t cols = 32;
  while (cols < (uint32_t)(P.nacc * P.ncols * (x3 ? 2 : 1))) cols <<= 1;
  if (cols > 512) return pl;
  P.tmem_cols = cols;
  P.in_ld2 = Cin * 2;
  pl.grid = units * P.G;
  pl.smem = 1024 + (size_t)P.nstages * P.stage_bytes + 256;
  pl.x3 = x3;
  pl.part_floats = (int64_t)pl.grid * P.nacc * P.prow * P.ncols;
  pl.ok = true;
  return pl;
}

}  // namespace

bool wgrad_tc_eligible(int N, int H, int W, int Cin, int Cout, int k, int stride, int passes) {
  return plan(N, H, W, Cin, Cout, k, stride, passes).ok;
}

int64_t wgrad_tc_workspace_bytes(int N, int H, int W, int Cin, int Cout, int k, int stride, int passes) {
  const WgPlan pl = plan(N, H, W, Cin, Cout, k, stride, passes);
  return pl.ok ? pl.part_floats * 4 : 0;
}

int launch_wgrad_tc(const float* x, const float* dy, float* dw, int N, int H, int W, int Cin, int Cout, int k, int stride,
                    int passes, void* workspace, cudaStream_t st) {
  WgPlan pl = plan(N, H, W, Cin, Cout, k, stride, passes);
  if (!pl.ok) return fail(LEDB200_EINVAL, "wgrad_tc: shape not eligible");
  WgParams& P = pl.P;
  P.part = reinterpret_cast<float*>(workspace);
  const int Ho = (H + 2 * (k / 2) - k) / stride + 1, Wo = (W + 2 * (k / 2) - k) / stride + 1;
  CUtensorMap tmX, tmY;
  int rc;
  const uint64_t ld = (uint64_t)Cin * 2;                 // 2-byte units per pixel
  if (P.k1 && !P.s2) {
    const uint64_t dims[4] = {ld, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t str[3] = {ld * 2, (uint64_t)W * ld * 2, (uint64_t)H * W * ld * 2};
    const uint32_t box[4] = {64, TW, TH, 1};
    rc = tc_encode_tiled(&tmX, x, 4, dims, str, box, 1064);
  } else if (P.k1) {
    const uint64_t dims[5] = {2 * ld, (uint64_t)W / 2, 2, (uint64_t)H / 2, (uint64_t)N};
    const uint64_t str[4] = {2 * ld * 2, (uint64_t)W * ld * 2, 2 * (uint64_t)W * ld * 2, (uint64_t)H * W * ld * 2};
    const uint32_t box[5] = {64, TW, 1, TH, 1};
    rc = tc_encode_tiled(&tmX, x, 5, dims, str, box, 1064);
  } else
  if (!P.s2) {
    const uint64_t dims[4] = {ld, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t str[3] = {ld * 2, (uint64_t)W * ld * 2, (uint64_t)H * W * ld * 2};
    const uint32_t box[4] = {64, (uint32_t)P.slab_w, (uint32_t)(TH + 2), 1};
    rc = tc_encode_tiled(&tmX, x, 4, dims, str, box, 1064);
  } else {
    const uint64_t dims[5] = {2 * ld, (uint64_t)W / 2, 2, (uint64_t)H / 2, (uint64_t)N};
    const uint64_t str[4] = {2 * ld * 2, (uint64_t)W * ld * 2, 2 * (uint64_t)W * ld * 2, (uint64_t)H * W * ld * 2};
    const uint32_t box[5] = {64, (uint32_t)P.slab_w, 1, (uint32_t)(TH + 1), 1};
    rc = tc_encode_tiled(&tmX, x, 5, dims, str, box, 1064);
  }
  if (rc) return rc;
  {
    const uint64_t ldy = (uint64_t)Cout * 2;
    const uint64_t dims[4] = {ldy, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)N};
    const uint64_t str[3] = {ldy * 2, (uint64_t)Wo * ldy * 2, (uint64_t)Ho * Wo * ldy * 2};
    const uint32_t box[4] = {64, TW, TH, 1};
    rc = tc_encode_tiled(&tmY, dy, 4, dims, str, box, 1064);
  }
  if (rc) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    const void* fns[6] = {(const void*)wgrad_tc_kernel<false, false>, (const void*)wgrad_tc_kernel<true, false>,
                          (const void*)wgrad_tc_kernel<false, true>, (const void*)wgrad_tc_kernel<true, true>,
                          (const void*)wgrad_tc_kernel<false, false, true>, (const void*)wgrad_tc_kernel<true, false, true>};
    for (const void* f : fns) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
      if (e != cudaSuccess) attr_err = e;
    }
  });
  if (attr_err != cudaSuccess) return fail(LEDB200_ECUDA, std::string("wgrad_tc: cudaFuncSetAttribute: ") + cudaGetErrorString(attr_err));
  const int nthr = pl.x3 ? kThreads + 128 : kThreads;
  if (P.k1) {
    if (P.s2) wgrad_tc_kernel<true, false, true><<<pl.grid, nthr, pl.smem, st>>>(tmX, tmY, P);
    else wgrad_tc_kernel<false, false, true><<<pl.grid, nthr, pl.smem, st>>>(tmX, tmY, P);
    LEDB_LAUNCH_OK("wgrad_tc_kernel");
    const int64_t total1 = (int64_t)Cout * Cin;
    wgrad_tc_sum1_kernel<<<(int)std::min<int64_t>(ceil_div64(total1, 256), 148 * 8), 256, 0, st>>>(P.part, dw, Cin, Cout, P.ncc, P.nco,
                                                                                              P.units_ci, P.G);
    LEDB_LAUNCH_OK("wgrad_tc_sum1_kernel");
    return LEDB200_OK;
  }
  if (pl.x3) {
    if (P.s2) wgrad_tc_kernel<true, true><<<pl.grid, nthr, pl.smem, st>>>(tmX, tmY, P);
    else wgrad_tc_kernel<false, true><<<pl.grid, nthr, pl.smem, st>>>(tmX, tmY, P);
  } else {
    if (P.s2) wgrad_tc_kernel<true, false><<<pl.grid, nthr, pl.smem, st>>>(tmX, tmY, P);
    else wgrad_tc_kernel<false, false><<<pl.grid, nthr, pl.smem, st>>>(tmX, tmY, P);
  }
  LEDB_LAUNCH_OK("wgrad_tc_kernel");
  const int64_t total = (int64_t)Cout * Cin * 9;
  wgrad_tc_sum_kernel<<<(int)std::min<int64_t>(ceil_div64(total, 256), 148 * 8), 256, 0, st>>>(P.part, dw, Cin, Cout, P.ncc, P.nco,
                                                                                           P.units_ci, P.G, P.s2);
  LEDB_LAUNCH_OK("wgrad_tc_sum_kernel");
  return LEDB200_OK;
}

}  // namespace ledb
