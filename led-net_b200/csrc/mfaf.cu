// MFAF gate (multi-scale attentional feature fusion) as three small kernels.
//
// Replaces Muti_AFF.forward (mmseg/models/classification/model_utils.py:410-429), eval mode:
//   xa  = x + residual
//   att = local(xa) + global(avgpool_1(xa)) + sum_{L in 4,8,16} nearest_up(ctx_L(adaptive_avgpool_LxL(xa)))
//   out = 2 x sigmoid(att) + 2 residual (1 - sigmoid(att))
// where every attention path is conv1x1(C -> CI, bias) -> BN -> ReLU -> conv1x1(CI -> C, bias) -> BN
// (model_utils.py:364-406).  The reference launches ~40 kernels and materialises xa, five attention maps and
// three up-sampled context maps; here x and residual are read twice (the pooled statistics are a global
// dependency) and the output written once:
//   1. mfaf_segsum_kernel  : sums of xa over the elementary cells cut out by ALL pooling-bin boundaries of the
//                            four pooled grids (every adaptive-pool bin, overlapping or not, is a union of cells);
//                            deterministic (no atomics): one thread per (x-cell, 8 channels), one CTA per y-cell.
//   2. mfaf_context_kernel : per (image, bin): mean -> the path's two 1x1 convs on one vector -> ctx table.
//   3. mfaf_gate_kernel    : per pixel: local path (C x CI + CI x C MACs, weights in shared memory) + the four
//                            context lookups (ATen's nearest index) -> sigmoid -> blend.
// Adaptive pooling bins follow ATen (start = floor(i*H/L), end = ceil((i+1)*H/L)); nearest up-sampling follows
// ATen's nearest_idx (identity / exact x2 shortcuts, else floorf(dst * (float)in / out) clamped).
#include "kernels.h"

namespace ledb {
namespace {

constexpr int MFAF_MAXCUT = 64;         // <= 2 * (16 + 8 + 4 + 1) + 1 distinct cut points per axis
constexpr int MFAF_BINS = 16 + 64 + 256 + 1;
constexpr int MFAF_PATHS = 5;           // local, ctx4, ctx8, ctx16, global

struct MfafArgs {
  const void* x;
  const void* res;
  void* out;
  const float* p;        // packed parameters: per path w1[CI][C], a1[CI], b1[CI], w2[C][CI], a2[C], b2[C]
  float* segsum;         // [N][ncy][ncx][C]
  float* ctx;            // [N][MFAF_BINS][C]
  int N, H, W, C, CI;
  int ncx, ncy;          // number of cells per axis
  short cutx[MFAF_MAXCUT], cuty[MFAF_MAXCUT];   // cell c spans [cut[c], cut[c+1])
};

__host__ __device__ inline int path_floats(int C, int CI) { return 2 * C * CI + 2 * CI + 2 * C; }
__host__ __device__ inline int bin_start(int i, int size, int L) { return (i * size) / L; }
__host__ __device__ inline int bin_end(int i, int size, int L) { return ((i + 1) * size + L - 1) / L; }

// ATen nearest_idx (aten/src/ATen/native/cpu/UpSampleKernel.cpp): output index -> input index
__device__ __forceinline__ int nearest_idx(int o, int in_size, int out_size) {
  if (out_size == in_size) return o;
  if (out_size == 2 * in_size) return o >> 1;
  const float scale = (float)in_size / (float)out_size;
  const int i = (int)floorf((float)o * scale);
  return i < in_size - 1 ? i : in_size - 1;
}

template <typename T>
__global__ void __launch_bounds__(256) mfaf_segsum_kernel(MfafArgs a) {
  const int cy = blockIdx.x, n = blockIdx.y;
  const int y0 = a.cuty[cy], y1 = a.cuty[cy + 1];
  const int cgs = a.C / 8;
  const T* x = reinterpret_cast<const T*>(a.x) + (int64_t)n * a.H * a.W * a.C;
  const T* r = reinterpret_cast<const T*>(a.res) + (int64_t)n * a.H * a.W * a.C;
  for (int item = threadIdx.x; item < a.ncx * cgs; item += blockDim.x) {
    const int cg = item % cgs, cx = item / cgs;
    const int x0 = a.cutx[cx], x1 = a.cutx[cx + 1];
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int y = y0; y < y1; ++y)
      for (int xx = x0; xx < x1; ++xx) {
        const int64_t off = ((int64_t)y * a.W + xx) * a.C + cg * 8;
        float u[8], v[8];
        load8(x + off, u);
        load8(r + off, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += u[q] + v[q];
      }
    store8(a.segsum + (((int64_t)n * a.ncy + cy) * a.ncx + cx) * a.C + cg * 8, acc);
  }
}

__global__ void __launch_bounds__(128) mfaf_context_kernel(MfafArgs a) {
  extern __shared__ float sm[];
  float* mean = sm;            // [C]
  float* hid = sm + a.C;       // [CI]
  const int bin = blockIdx.x, n = blockIdx.y;
  int L, b, path;
  if (bin < 16) { L = 4; b = bin; path = 1; }
  else if (bin < 80) { L = 8; b = bin - 16; path = 2; }
  else if (bin < 336) { L = 16; b = bin - 80; path = 3; }
  else { L = 1; b = 0; path = 4; }
  const int by = b / L, bx = b % L;
  const int ys = bin_start(by, a.H, L), ye = bin_end(by, a.H, L);
  const int xs = bin_start(bx, a.W, L), xe = bin_end(bx, a.W, L);
  int cy0 = 0, cy1 = 0, cx0 = 0, cx1 = 0;
  for (int i = 0; i <= a.ncy; ++i) { if (a.cuty[i] == ys) cy0 = i; if (a.cuty[i] == ye) cy1 = i; }
  for (int i = 0; i <= a.ncx; ++i) { if (a.cutx[i] == xs) cx0 = i; if (a.cutx[i] == xe) cx1 = i; }
  const float inv = 1.f / (float)((ye - ys) * (xe - xs));
  const float* seg = a.segsum + (int64_t)n * a.ncy * a.ncx * a.C;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float s = 0.f;
    for (int cy = cy0; cy < cy1; ++cy)
      for (int cx = cx0; cx < cx1; ++cx) s += seg[((int64_t)cy * a.ncx + cx) * a.C + c];
    mean[c] = s * inv;
  }
  __syncthreads();
  const float* P = a.p + (int64_t)path * path_floats(a.C, a.CI);
  const float *w1 = P, *a1 = w1 + a.CI * a.C, *b1 = a1 + a.CI, *w2 = b1 + a.CI, *a2 = w2 + a.C * a.CI, *b2 = a2 + a.C;
  for (int j = threadIdx.x; j < a.CI; j += blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < a.C; ++c) s = fmaf(w1[j * a.C + c], mean[c], s);
    hid[j] = fmaxf(fmaf(s, a1[j], b1[j]), 0.f);
  }
  __syncthreads();
  float* dst = a.ctx + ((int64_t)n * MFAF_BINS + bin) * a.C;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < a.CI; ++j) s = fmaf(w2[c * a.CI + j], hid[j], s);
    dst[c] = fmaf(s, a2[c], b2[c]);
  }
}

template <typename T, int CI>
__global__ void __launch_bounds__(128) mfaf_gate_kernel(MfafArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int C = a.C;
  float* w1t = sm;                 // [C][CI]  (transposed: input channel major)
  float* w2 = w1t + C * CI;        // [C][CI]
  float* a1 = w2 + C * CI;         // [CI]
  float* b1 = a1 + CI;             // [CI]
  float* a2 = b1 + CI;             // [C]
  float* b2 = a2 + C;              // [C]
  {
    const float* P = a.p;          // path 0 = local_att
    const float *g1 = P, *ga1 = g1 + CI * C, *gb1 = ga1 + CI, *g2 = gb1 + CI, *ga2 = g2 + C * CI, *gb2 = ga2 + C;
    for (int i = threadIdx.x; i < C * CI; i += blockDim.x) {
      w1t[i] = g1[(i % CI) * C + i / CI];
      w2[i] = g2[i];
    }
    for (int i = threadIdx.x; i < CI; i += blockDim.x) { a1[i] = ga1[i]; b1[i] = gb1[i]; }
    for (int i = threadIdx.x; i < C; i += blockDim.x) { a2[i] = ga2[i]; b2[i] = gb2[i]; }
  }
  __syncthreads();
  const int64_t npix = (int64_t)a.N * a.H * a.W;
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int xw = (int)(pix % a.W);
  const int yh = (int)((pix / a.W) % a.H);
  const int n = (int)(pix / ((int64_t)a.W * a.H));
  const T* x = reinterpret_cast<const T*>(a.x) + pix * C;
  const T* r = reinterpret_cast<const T*>(a.res) + pix * C;
  T* out = reinterpret_cast<T*>(a.out) + pix * C;
  // ---- local path, first 1x1: hidden = relu(a1 * (W1 xa) + b1)
  float h[CI];
#pragma unroll
  for (int j = 0; j < CI; ++j) h[j] = 0.f;
  for (int c0 = 0; c0 < C; c0 += 8) {
    float u[8], v[8];
    load8(x + c0, u);
    load8(r + c0, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float xa = u[q] + v[q];
      const float4* w = reinterpret_cast<const float4*>(w1t + (c0 + q) * CI);
#pragma unroll
      for (int j = 0; j < CI / 4; ++j) {
        const float4 w4 = w[j];
        h[4 * j + 0] = fmaf(xa, w4.x, h[4 * j + 0]); h[4 * j + 1] = fmaf(xa, w4.y, h[4 * j + 1]);
        h[4 * j + 2] = fmaf(xa, w4.z, h[4 * j + 2]); h[4 * j + 3] = fmaf(xa, w4.w, h[4 * j + 3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CI; ++j) h[j] = fmaxf(fmaf(h[j], a1[j], b1[j]), 0.f);
  // ---- context rows of this pixel (ATen nearest up-sampling of the 4x4 / 8x8 / 16x16 grids) + global
  const float* cb = a.ctx + (int64_t)n * MFAF_BINS * C;
  const float* c4 = cb + (nearest_idx(yh, 4, a.H) * 4 + nearest_idx(xw, 4, a.W)) * C;
  const float* c8 = cb + (16 + nearest_idx(yh, 8, a.H) * 8 + nearest_idx(xw, 8, a.W)) * C;
  const float* c16 = cb + (80 + nearest_idx(yh, 16, a.H) * 16 + nearest_idx(xw, 16, a.W)) * C;
  const float* cg = cb + 336 * C;
  // ---- second 1x1 + contexts -> sigmoid -> blend
  for (int c0 = 0; c0 < C; c0 += 8) {
    float u[8], v[8], o[8];
    load8(x + c0, u);
    load8(r + c0, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = c0 + q;
      const float4* w = reinterpret_cast<const float4*>(w2 + c * CI);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < CI / 4; ++j) {
        const float4 w4 = w[j];
        s = fmaf(h[4 * j + 0], w4.x, fmaf(h[4 * j + 1], w4.y, fmaf(h[4 * j + 2], w4.z, fmaf(h[4 * j + 3], w4.w, s))));
      }
      const float att = fmaf(s, a2[c], b2[c]) + __ldg(cg + c) + __ldg(c4 + c) + __ldg(c8 + c) + __ldg(c16 + c);
      const float wei = 1.f / (1.f + __expf(-att));
      o[q] = 2.f * u[q] * wei + 2.f * v[q] * (1.f - wei);
    }
    store8(out + c0, o);
  }
}

// distinct, sorted bin boundaries of the 1 / 4 / 8 / 16 grids over `size`
int make_cuts(int size, short* cuts) {
  int n = 0;
  short tmp[2 * (16 + 8 + 4 + 1)];
  for (int L : {1, 4, 8, 16})
    for (int i = 0; i < L; ++i) { tmp[n++] = (short)bin_start(i, size, L); tmp[n++] = (short)bin_end(i, size, L); }
  // insertion sort + unique (n = 58)
  for (int i = 1; i < n; ++i) { short v = tmp[i]; int j = i - 1; while (j >= 0 && tmp[j] > v) { tmp[j + 1] = tmp[j]; --j; } tmp[j + 1] = v; }
  int m = 0;
  for (int i = 0; i < n; ++i) if (m == 0 || cuts[m - 1] != tmp[i]) cuts[m++] = tmp[i];
  return m - 1;   // number of cells
}

}  // namespace
}  // namespace ledb

using namespace ledb;

extern "C" {

int64_t ledb200_mfaf_param_floats(int32_t C, int32_t CI) { return (int64_t)MFAF_PATHS * path_floats(C, CI); }

int64_t ledb200_mfaf_workspace_bytes(int32_t N, int32_t C) {
  return (int64_t)N * ((int64_t)(MFAF_MAXCUT - 1) * (MFAF_MAXCUT - 1) + MFAF_BINS) * C * (int64_t)sizeof(float);
}

int ledb200_mfaf_forward(const void* x, const void* residual, void* out, int32_t dtype, int32_t N, int32_t H,
                         int32_t W, int32_t C, int32_t CI, const float* params, void* workspace, void* stream) {
  if (!x || !residual || !out || !params || !workspace) return fail(LEDB200_EINVAL, "mfaf: null buffer");
  if (dtype != LEDB200_F32 && dtype != LEDB200_BF16) return fail(LEDB200_EINVAL, "mfaf: dtype must be F32 or BF16");
  if (N < 1 || H < 1 || W < 1) return fail(LEDB200_EINVAL, "mfaf: empty input");
  if (H > 32767 || W > 32767) return fail(LEDB200_EINVAL, "mfaf: H and W must be below 32768");
  if (C % 8 || C < 8 || C > 256) return fail(LEDB200_EINVAL, "mfaf: channels must be a multiple of 8 in [8, 256]");
  if (CI != 8 && CI != 16 && CI != 32 && CI != 64)
    return fail(LEDB200_EINVAL, "mfaf: channels // r must be 8, 16, 32 or 64");
  MfafArgs a;
  a.x = x; a.res = residual; a.out = out; a.p = params; a.N = N; a.H = H; a.W = W; a.C = C; a.CI = CI;
  a.ncx = make_cuts(W, a.cutx);
  a.ncy = make_cuts(H, a.cuty);
  a.segsum = reinterpret_cast<float*>(workspace);
  a.ctx = a.segsum + (int64_t)N * (MFAF_MAXCUT - 1) * (MFAF_MAXCUT - 1) * C;
  cudaStream_t st = (cudaStream_t)stream;
  const bool bf = dtype == LEDB200_BF16;
  if (bf) mfaf_segsum_kernel<__nv_bfloat16><<<dim3(a.ncy, N), 256, 0, st>>>(a);
  else    mfaf_segsum_kernel<float><<<dim3(a.ncy, N), 256, 0, st>>>(a);
  LEDB_LAUNCH_OK("mfaf_segsum_kernel");
  mfaf_context_kernel<<<dim3(MFAF_BINS, N), 128, (C + CI) * sizeof(float), st>>>(a);
  LEDB_LAUNCH_OK("mfaf_context_kernel");
  const size_t smem = sizeof(float) * ((size_t)2 * C * CI + 2 * CI + 2 * C);
  const int64_t npix = (int64_t)N * H * W;
  const unsigned grid = (unsigned)ceil_div64(npix, 128);
#define MFAF_GATE(TT, CC)                                                                                         \
  do {                                                                                                            \
    LEDB_CUDA_OK(cudaFuncSetAttribute(mfaf_gate_kernel<TT, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mfaf_gate_kernel<TT, CC><<<grid, 128, smem, st>>>(a);                                                         \
  } while (0)
  if (bf) {
    if (CI == 8) MFAF_GATE(__nv_bfloat16, 8); else if (CI == 16) MFAF_GATE(__nv_bfloat16, 16);
    else if (CI == 32) MFAF_GATE(__nv_bfloat16, 32); else MFAF_GATE(__nv_bfloat16, 64);
  } else {
    if (CI == 8) MFAF_GATE(float, 8); else if (CI == 16) MFAF_GATE(float, 16);
    else if (CI == 32) MFAF_GATE(float, 32); else MFAF_GATE(float, 64);
  }
#undef MFAF_GATE
  LEDB_LAUNCH_OK("mfaf_gate_kernel");
  return LEDB200_OK;
}

}  // extern "C"
