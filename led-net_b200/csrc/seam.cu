// SEAM edge gate, eval mode.  Follows the inline edge path of the authors' speed
// prototype tools/speed/ddrnet_speed.py (parameters :88-113, edge map :282-338, gate :388-389):
//   e = minmax_normalise(BN(conv3x3 C->1 (x)))  over the whole tensor (batch included, :24-37);
//   b_s = [clamp(laplacian_stride_s(e), 0) > t], s = 1, 2, 4 (strided maps nearest-upsampled, ATen index rule);
//   m = [0.6 b_1 + 0.3 b_2 + 0.1 b_4 > t];  out = BN(conv3x3 1->C (m)) * x_s + x_s.
// Three kernels: (1) the 1-channel edge response + a global min / max (order-preserving integer atomics),
// (2) the binary mask from the three Laplacians (each evaluated only where it is sampled), (3) the gate: nine mask
// taps x C weights per pixel, x_s read once and the result written once.
#include "kernels.h"

namespace ledb {
namespace {

struct SeamArgs {
  const void* x;      // [N,H,W,C] edge source
  const void* xs;     // [N,H,W,C] gated tensor
  void* out;          // [N,H,W,C]
  const float* p;     // w1[9][C], a1, b1, 6 pad floats, w2[9][C], a2[C], b2[C]   (BN folded: y = a * conv + b)
  float* e;           // [N,H,W]
  uint8_t* mask;      // [N,H,W]
  unsigned* mm;       // [2] encoded min, max
  int N, H, W, C;
  float thr;
};

// order-preserving float <-> uint (atomicMin / atomicMax on floats of either sign)
__device__ __forceinline__ unsigned f2ord(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__device__ __forceinline__ int nearest_idx(int o, int in_size, int out_size) {   // ATen nearest_idx
  if (out_size == in_size) return o;
  if (out_size == 2 * in_size) return o >> 1;
  const float scale = (float)in_size / (float)out_size;
  const int i = (int)floorf((float)o * scale);
  return i < in_size - 1 ? i : in_size - 1;
}

__global__ void seam_init_kernel(unsigned* mm) { mm[0] = 0xffffffffu; mm[1] = 0u; }

template <typename T>
__global__ void __launch_bounds__(256) seam_edge_kernel(SeamArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t npix = (int64_t)a.N * a.H * a.W;
  float v = 0.f;
  const bool ok = idx < npix;
  if (ok) {
    const int x = (int)(idx % a.W), y = (int)((idx / a.W) % a.H);
    const int n = (int)(idx / ((int64_t)a.W * a.H));
    const T* src = reinterpret_cast<const T*>(a.x) + (int64_t)n * a.H * a.W * a.C;
    float acc = 0.f;
    for (int kh = 0; kh < 3; ++kh) {
      const int yy = y + kh - 1;
      if (yy < 0 || yy >= a.H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int xx = x + kw - 1;
        if (xx < 0 || xx >= a.W) continue;
        const T* px = src + ((int64_t)yy * a.W + xx) * a.C;
        const float* w = a.p + (kh * 3 + kw) * a.C;
        for (int c = 0; c < a.C; c += 8) {
          float t[8], u[8];
          load8(px + c, t);
          load8(w + c, u);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc = fmaf(t[q], u[q], acc);
        }
      }
    }
    v = fmaf(acc, a.p[9 * a.C], a.p[9 * a.C + 1]);
    a.e[idx] = v;
  }
  // block min / max, one atomic pair per warp
  float mn = ok ? v : INFINITY, mx = ok ? v : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0 && mn <= mx) { atomicMin(&a.mm[0], f2ord(mn)); atomicMax(&a.mm[1], f2ord(mx)); }
}

// clamp(Laplacian at centre (cy, cx) of the normalised map, zero padded) > thr
__device__ __forceinline__ bool lap_gt(const float* e, int H, int W, int cy, int cx, float mn, float den, float thr) {
  float acc = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int yy = cy + kh - 1, xx = cx + kw - 1;
      const float v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (e[(int64_t)yy * W + xx] - mn) / den : 0.f;
      acc = fmaf((kh == 1 && kw == 1) ? 8.f : -1.f, v, acc);
    }
  return fmaxf(acc, 0.f) > thr;
}

__global__ void __launch_bounds__(256) seam_mask_kernel(SeamArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)a.N * a.H * a.W) return;
  const int x = (int)(idx % a.W), y = (int)((idx / a.W) % a.H);
  const int n = (int)(idx / ((int64_t)a.W * a.H));
  const float mn = ord2f(a.mm[0]), den = ord2f(a.mm[1]) - mn;
  const float* e = a.e + (int64_t)n * a.H * a.W;
  const int H2 = (a.H - 1) / 2 + 1, W2 = (a.W - 1) / 2 + 1, H4 = (a.H - 1) / 4 + 1, W4 = (a.W - 1) / 4 + 1;
  const float b1 = lap_gt(e, a.H, a.W, y, x, mn, den, a.thr) ? 1.f : 0.f;
  const float b2 = lap_gt(e, a.H, a.W, 2 * nearest_idx(y, H2, a.H), 2 * nearest_idx(x, W2, a.W), mn, den, a.thr) ? 1.f : 0.f;
  const float b4 = lap_gt(e, a.H, a.W, 4 * nearest_idx(y, H4, a.H), 4 * nearest_idx(x, W4, a.W), mn, den, a.thr) ? 1.f : 0.f;
  const float pyr = fmaf(0.1f, b4, fmaf(0.3f, b2, 0.6f * b1));        // fusion_kernel = (6/10, 3/10, 1/10)
  a.mask[idx] = pyr > a.thr ? 1 : 0;
}

template <typename T>
__global__ void __launch_bounds__(256) seam_gate_kernel(SeamArgs a) {
  const int cgs = a.C / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)a.N * a.H * a.W * cgs) return;
  const int cg = (int)(idx % cgs);
  const int64_t pix = idx / cgs;
  const int x = (int)(pix % a.W), y = (int)((pix / a.W) % a.H);
  const int n = (int)(pix / ((int64_t)a.W * a.H));
  const uint8_t* m = a.mask + (int64_t)n * a.H * a.W;
  const float* w2 = a.p + 9 * a.C + 8;
  float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int yy = y + kh - 1, xx = x + kw - 1;
      if (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W && m[(int64_t)yy * a.W + xx]) {
        float w[8];
        load8(w2 + (kh * 3 + kw) * a.C + cg * 8, w);
#pragma unroll
        for (int q = 0; q < 8; ++q) g[q] += w[q];
      }
    }
  const float* a2 = w2 + 9 * a.C;
  const float* b2 = a2 + a.C;
  float xs[8], o[8];
  load8(reinterpret_cast<const T*>(a.xs) + pix * a.C + cg * 8, xs);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float gate = fmaf(g[q], a2[cg * 8 + q], b2[cg * 8 + q]);
    o[q] = __fadd_rn(__fmul_rn(gate, xs[q]), xs[q]);            // result = conv_2(m) * x_s ; x_s = result + x_s
  }
  store8(reinterpret_cast<T*>(a.out) + pix * a.C + cg * 8, o);
}

}  // namespace
}  // namespace ledb

using namespace ledb;

extern "C" {

int64_t ledb200_seam_param_floats(int32_t C) { return 9 * (int64_t)C + 8 + 9 * (int64_t)C + 2 * (int64_t)C; }
int64_t ledb200_seam_workspace_bytes(int32_t N, int32_t H, int32_t W) { return (int64_t)N * H * W * 5 + 64; }

int ledb200_seam_forward(const void* x, const void* x_s, void* out, int32_t dtype, int32_t N, int32_t H, int32_t W,
                         int32_t C, float threshold, const float* params, void* workspace, void* stream) {
  if (!x || !x_s || !out || !params || !workspace) return fail(LEDB200_EINVAL, "seam: null buffer");
  if (dtype != LEDB200_F32 && dtype != LEDB200_BF16) return fail(LEDB200_EINVAL, "seam: dtype must be F32 or BF16");
  if (N < 1 || H < 1 || W < 1) return fail(LEDB200_EINVAL, "seam: empty input");
  if (C < 8 || C % 8) return fail(LEDB200_EINVAL, "seam: channels must be a multiple of 8");
  SeamArgs a;
  const int64_t npix = (int64_t)N * H * W;
  a.x = x; a.xs = x_s; a.out = out; a.p = params; a.N = N; a.H = H; a.W = W; a.C = C; a.thr = threshold;
  a.mm = reinterpret_cast<unsigned*>(workspace);
  a.e = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 64);
  a.mask = reinterpret_cast<uint8_t*>(a.e + npix);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned gp = (unsigned)ceil_div64(npix, 256), gg = (unsigned)ceil_div64(npix * (C / 8), 256);
  seam_init_kernel<<<1, 1, 0, st>>>(a.mm);
  if (dtype == LEDB200_BF16) seam_edge_kernel<__nv_bfloat16><<<gp, 256, 0, st>>>(a);
  else seam_edge_kernel<float><<<gp, 256, 0, st>>>(a);
  LEDB_LAUNCH_OK("seam_edge_kernel");
  seam_mask_kernel<<<gp, 256, 0, st>>>(a);
  LEDB_LAUNCH_OK("seam_mask_kernel");
  if (dtype == LEDB200_BF16) seam_gate_kernel<__nv_bfloat16><<<gg, 256, 0, st>>>(a);
  else seam_gate_kernel<float><<<gg, 256, 0, st>>>(a);
  LEDB_LAUNCH_OK("seam_gate_kernel");
  return LEDB200_OK;
}

}  // extern "C"
