// HBM-bound elementwise / resampling kernels (north_star kernel 3 and DAPPM glue).
//
//  * upsample_add : `x_s += resize(comp_c, size=out_size, mode='bilinear', align_corners=False)`
//                   (mmseg/models/backbones/ddrnet.py:195-199, 208-212, 218-224) and the DAPPM
//                   `F.interpolate(scale_i(x)) + feats[i-1]` (mmseg/models/utils/ppm.py:124-127),
//                   with the consumer's ReLU / pre-activation BN+ReLU fused so each branch is
//                   read once and the sum is written once.
//  * avgpool_bnrelu : nn.AvgPool2d(5,2,2)/(9,4,4)/(17,8,8) and AdaptiveAvgPool2d(1) of DAPPM
//                   (ppm.py:66-90) + the following ConvModule's BN+ReLU prologue.
//  * affine_relu  : per-channel BN+ReLU prologue feeding two pre-activation 1x1 convs
//                   (DAPPM scales[0] and shortcut, ppm.py:57-63, 110-117).
//  * nchw<->nhwc  : the boundary to the reference's NCHW fp32 tensors.
// All NHWC, 8 channels (16 B bf16 / 32 B fp32) per thread, grid sized to 148 SMs x 8 CTAs.
#include "kernels.h"

namespace ledb {
namespace {

constexpr int kThreads = 256;
inline int grid_for(int64_t work) {
  int64_t b = ceil_div64(work, kThreads);
  const int64_t cap = 148 * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) upsample_add_kernel(UpAddArgs a, float sh, float sw) {
  const int cg = a.C / 8;
  const int64_t total = (int64_t)a.N * a.H * a.W * cg;
  const T* base = reinterpret_cast<const T*>(a.base);
  const T* src = reinterpret_cast<const T*>(a.src);
  T* out = reinterpret_cast<T*>(a.out);
  T* out2 = reinterpret_cast<T*>(a.out2);
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * kThreads) {
    const int g = (int)(i % cg);
    int64_t p = i / cg;
    const int x = (int)(p % a.W);
    const int y = (int)((p / a.W) % a.H);
    const int n = (int)(p / ((int64_t)a.W * a.H));
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilinear_coord(y, sh, a.h, y0, y1, ly0, ly1);
    bilinear_coord(x, sw, a.w, x0, x1, lx0, lx1);
    const T* s = src + ((int64_t)n * a.h * a.w) * a.C + g * 8;
    float v00[8], v01[8], v10[8], v11[8], v[8];
    load8(s + ((int64_t)y0 * a.w + x0) * a.C, v00);
    load8(s + ((int64_t)y0 * a.w + x1) * a.C, v01);
    load8(s + ((int64_t)y1 * a.w + x0) * a.C, v10);
    load8(s + ((int64_t)y1 * a.w + x1) * a.C, v11);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float r0 = fmaf(v01[c], lx1, v00[c] * lx0);
      const float r1 = fmaf(v11[c], lx1, v10[c] * lx0);
      v[c] = fmaf(r1, ly1, r0 * ly0);
    }
    if (base) {
      float b[8];
      load8(base + p * a.C + g * 8, b);
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] += b[c];
    }
    if (a.relu) {
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = fmaxf(v[c], 0.f);
    }
    if (out) store8(out + p * a.out_ld + g * 8, v);
    if (out2) {
      float o[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int ch = g * 8 + c;
        const float t = a.o2_scale ? fmaf(v[c], a.o2_scale[ch], a.o2_shift[ch]) : v[c];
        o[c] = fmaxf(t, 0.f);
      }
      store8(out2 + p * a.out2_ld + g * 8, o);
    }
  }
}


// bf16 fast path: 16 channels (one 32 B sector) per thread through 256-bit loads / stores, 32-bit index
// arithmetic (the generic kernel above spends most of its issue slots on 64-bit div/mod and the bilinear
// coordinates, which are amortised over twice the channels here).  ncu r1b: generic kernel issue-bound (76 %)
// at 35-40 % of the HBM roofline.
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* f) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

__global__ void __launch_bounds__(kThreads) upsample_add16_kernel(UpAddArgs a, float sh, float sw, uint32_t total) {
  const uint32_t cg = (uint32_t)a.C / 16;
  const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(a.base);
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(a.src);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
  __nv_bfloat16* out2 = reinterpret_cast<__nv_bfloat16*>(a.out2);
  for (uint32_t i = blockIdx.x * (uint32_t)kThreads + threadIdx.x; i < total; i += gridDim.x * (uint32_t)kThreads) {
    const uint32_t g = i % cg, p = i / cg;
    const uint32_t x = p % (uint32_t)a.W, t = p / (uint32_t)a.W;
    const uint32_t y = t % (uint32_t)a.H, n = t / (uint32_t)a.H;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilinear_coord((int)y, sh, a.h, y0, y1, ly0, ly1);
    bilinear_coord((int)x, sw, a.w, x0, x1, lx0, lx1);
    const __nv_bfloat16* s = src + ((int64_t)n * a.h * a.w) * a.C + g * 16;
    uint4 q[4][2];
    ldg256(s + (int64_t)(y0 * a.w + x0) * a.C, q[0][0], q[0][1]);
    ldg256(s + (int64_t)(y0 * a.w + x1) * a.C, q[1][0], q[1][1]);
    ldg256(s + (int64_t)(y1 * a.w + x0) * a.C, q[2][0], q[2][1]);
    ldg256(s + (int64_t)(y1 * a.w + x1) * a.C, q[3][0], q[3][1]);
    uint4 bq[2];
    if (base) ldg256(base + (int64_t)p * a.C + g * 16, bq[0], bq[1]);
    uint4 o[2], o2[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float v00[8], v01[8], v10[8], v11[8], v[8];
      unpack_bf16x8(q[0][hh], v00); unpack_bf16x8(q[1][hh], v01);
      unpack_bf16x8(q[2][hh], v10); unpack_bf16x8(q[3][hh], v11);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float r0 = fmaf(v01[c], lx1, v00[c] * lx0);
        const float r1 = fmaf(v11[c], lx1, v10[c] * lx0);
        v[c] = fmaf(r1, ly1, r0 * ly0);
      }
      if (base) {
        float b[8];
        unpack_bf16x8(bq[hh], b);
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] += b[c];
      }
      if (a.relu) {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = fmaxf(v[c], 0.f);
      }
      o[hh] = pack_bf16x8(v);
      if (out2) {
        float w[8];
        if (a.o2_scale) {                      // per-channel affine as four 16 B loads (arena buffers: 16 B aligned)
          const float4* sc = reinterpret_cast<const float4*>(a.o2_scale + g * 16 + hh * 8);
          const float4* sf = reinterpret_cast<const float4*>(a.o2_shift + g * 16 + hh * 8);
          const float4 s0 = __ldg(sc), s1 = __ldg(sc + 1), b0 = __ldg(sf), b1 = __ldg(sf + 1);
          w[0] = fmaf(v[0], s0.x, b0.x); w[1] = fmaf(v[1], s0.y, b0.y); w[2] = fmaf(v[2], s0.z, b0.z); w[3] = fmaf(v[3], s0.w, b0.w);
          w[4] = fmaf(v[4], s1.x, b1.x); w[5] = fmaf(v[5], s1.y, b1.y); w[6] = fmaf(v[6], s1.z, b1.z); w[7] = fmaf(v[7], s1.w, b1.w);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) w[c] = v[c];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) w[c] = fmaxf(w[c], 0.f);
        o2[hh] = pack_bf16x8(w);
      }
    }
    if (out) stg256(out + (int64_t)p * a.out_ld + g * 16, o[0], o[1]);
    if (out2) stg256(out2 + (int64_t)p * a.out2_ld + g * 16, o2[0], o2[1]);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) avgpool_kernel(PoolArgs a) {
  const int cg = a.C / 8;
  const int64_t total = (int64_t)a.N * a.Ho * a.Wo * cg;
  const T* in = reinterpret_cast<const T*>(a.in);
  T* out = reinterpret_cast<T*>(a.out);
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * kThreads) {
    const int g = (int)(i % cg);
    int64_t p = i / cg;
    const int ox = (int)(p % a.Wo);
    const int oy = (int)((p / a.Wo) % a.Ho);
    const int n = (int)(p / ((int64_t)a.Wo * a.Ho));
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int y0, y1, x0, x1;
    float inv;
    if (a.k == 0) {   // global
      y0 = 0; y1 = a.H; x0 = 0; x1 = a.W;
      inv = 1.f / (float)(a.H * a.W);
    } else {          // count_include_pad=True: divisor is always k*k
      y0 = oy * a.s - a.p; y1 = y0 + a.k; x0 = ox * a.s - a.p; x1 = x0 + a.k;
      // PyTorch clips the window to the padded extent before counting (pool_size uses
      // min(hend, H+pad)); with k <= 2*pad+1 windows here never leave it.
      const int hend = min(y1, a.H + a.p), wend = min(x1, a.W + a.p);
      inv = 1.f / (float)((hend - y0) * (wend - x0));
      y0 = max(y0, 0); x0 = max(x0, 0); y1 = min(y1, a.H); x1 = min(x1, a.W);
    }
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        float v[8];
        load8(in + (((int64_t)n * a.H + y) * a.W + x) * a.C + g * 8, v);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] += v[c];
      }
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int ch = g * 8 + c;
      float v = acc[c] * inv;
      if (a.scale) v = fmaxf(fmaf(v, a.scale[ch], a.shift[ch]), 0.f);
      o[c] = v;
    }
    store8(out + p * a.C + g * 8, o);
  }
}


// Large windows (17x17 / global, DAPPM scales 3 and 4): one WARP per (output pixel, 8 channels), the lanes
// stride over the window and combine with shuffles.  The thread-per-output kernel above leaves 1-8 CTAs
// summing 289-512 positions serially (ncu r1b: 37 and 79 us for 2 MB of input).
template <typename T>
__global__ void __launch_bounds__(kThreads) avgpool_warp_kernel(PoolArgs a) {
  const int cg = a.C / 8;
  const int64_t total = (int64_t)a.N * a.Ho * a.Wo * cg;
  const T* in = reinterpret_cast<const T*>(a.in);
  T* out = reinterpret_cast<T*>(a.out);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * kThreads) >> 5;
  for (int64_t i = warp0; i < total; i += nwarps) {
    const int g = (int)(i % cg);
    const int64_t p = i / cg;
    const int ox = (int)(p % a.Wo);
    const int oy = (int)((p / a.Wo) % a.Ho);
    const int n = (int)(p / ((int64_t)a.Wo * a.Ho));
    int y0, y1, x0, x1;
    float inv;
    if (a.k == 0) {
      y0 = 0; y1 = a.H; x0 = 0; x1 = a.W;
      inv = 1.f / (float)(a.H * a.W);
    } else {
      y0 = oy * a.s - a.p; y1 = y0 + a.k; x0 = ox * a.s - a.p; x1 = x0 + a.k;
      const int hend = min(y1, a.H + a.p), wend = min(x1, a.W + a.p);
      inv = 1.f / (float)((hend - y0) * (wend - x0));
      y0 = max(y0, 0); x0 = max(x0, 0); y1 = min(y1, a.H); x1 = min(x1, a.W);
    }
    const int ww = x1 - x0, cnt = (y1 - y0) * ww;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = lane; j < cnt; j += 32) {
      const int y = y0 + j / ww, x = x0 + j % ww;
      float v[8];
      load8(in + (((int64_t)n * a.H + y) * a.W + x) * a.C + g * 8, v);
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] += v[c];
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
    if (lane == 0) {
      float o[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int ch = g * 8 + c;
        float v = acc[c] * inv;
        if (a.scale) v = fmaxf(fmaf(v, a.scale[ch], a.shift[ch]), 0.f);
        o[c] = v;
      }
      store8(out + p * a.C + g * 8, o);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) affine_relu_kernel(AffineArgs a) {
  const int cg = a.C / 8;
  const int64_t total = a.npix * cg;
  const T* in = reinterpret_cast<const T*>(a.in);
  T* oa = reinterpret_cast<T*>(a.out_a);
  T* ob = reinterpret_cast<T*>(a.out_b);
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * kThreads) {
    const int g = (int)(i % cg);
    float v[8], o[8];
    load8(in + i * 8, v);
    if (oa) {
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = fmaxf(fmaf(v[c], a.sa[g * 8 + c], a.ba[g * 8 + c]), 0.f);
      store8(oa + i * 8, o);
    }
    if (ob) {
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = fmaxf(fmaf(v[c], a.sb[g * 8 + c], a.bb[g * 8 + c]), 0.f);
      store8(ob + i * 8, o);
    }
  }
}

// NCHW fp32 -> NHWC T through a 32x32 smem transpose per (n, 32 pixels, 32 channels)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int C, int64_t HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const int64_t p = p0 + tx;
    tile[j][tx] = (c < C && p < HW) ? in[((int64_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int64_t p = p0 + j;
    const int c = c0 + tx;
    if (c < C && p < HW) out[((int64_t)n * HW + p) * C + c] = from_f32<T>(tile[tx][j]);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int C, int64_t HW, int ld) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int j = ty; j < 32; j += 8) {
    const int64_t p = p0 + j;
    const int c = c0 + tx;
    tile[j][tx] = (c < C && p < HW) ? to_f32(in[((int64_t)n * HW + p) * ld + c]) : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j;
    const int64_t p = p0 + tx;
    if (c < C && p < HW) out[((int64_t)n * C + c) * HW + p] = tile[tx][j];
  }
}

}  // namespace

int launch_upsample_add(const UpAddArgs& a, cudaStream_t st) {
  if (a.C % 8) return fail(LEDB200_EINVAL, "upsample_add: C must be a multiple of 8");
  const float sh = (float)a.h / (float)a.H, sw = (float)a.w / (float)a.W;
  const int64_t total = (int64_t)a.N * a.H * a.W * (a.C / 8);
  const bool al32 = a.C % 16 == 0 && (!a.out || a.out_ld % 16 == 0) && (!a.out2 || a.out2_ld % 16 == 0) &&
                    (((uintptr_t)a.base | (uintptr_t)a.src | (uintptr_t)a.out | (uintptr_t)a.out2) % 32 == 0) &&
                    (((uintptr_t)a.o2_scale | (uintptr_t)a.o2_shift) % 16 == 0);
  if (a.dtype == LEDB200_BF16 && al32 && total / 2 < (1ll << 31) && (int64_t)a.h * a.w * a.C < (1ll << 31)) {
    const uint32_t t16 = (uint32_t)(total / 2);
    upsample_add16_kernel<<<grid_for(t16), kThreads, 0, st>>>(a, sh, sw, t16);
  } else if (a.dtype == LEDB200_BF16)
    upsample_add_kernel<__nv_bfloat16><<<grid_for(total), kThreads, 0, st>>>(a, sh, sw);
  else
    upsample_add_kernel<float><<<grid_for(total), kThreads, 0, st>>>(a, sh, sw);
  LEDB_LAUNCH_OK("upsample_add_kernel");
  return LEDB200_OK;
}

int launch_avgpool_bnrelu(const PoolArgs& a, cudaStream_t st) {
  if (a.C % 8) return fail(LEDB200_EINVAL, "avgpool: C must be a multiple of 8");
  const int64_t total = (int64_t)a.N * a.Ho * a.Wo * (a.C / 8);
  if (a.k == 0 || a.k * a.k >= 200) {         // big windows (17x17, global): warp per output
    if (a.dtype == LEDB200_BF16) avgpool_warp_kernel<__nv_bfloat16><<<grid_for(total * 32), kThreads, 0, st>>>(a);
    else avgpool_warp_kernel<float><<<grid_for(total * 32), kThreads, 0, st>>>(a);
    LEDB_LAUNCH_OK("avgpool_warp_kernel");
    return LEDB200_OK;
  }
  if (a.dtype == LEDB200_BF16)
    avgpool_kernel<__nv_bfloat16><<<grid_for(total), kThreads, 0, st>>>(a);
  else
    avgpool_kernel<float><<<grid_for(total), kThreads, 0, st>>>(a);
  LEDB_LAUNCH_OK("avgpool_kernel");
  return LEDB200_OK;
}

int launch_affine_relu(const AffineArgs& a, cudaStream_t st) {
  if (a.C % 8) return fail(LEDB200_EINVAL, "affine_relu: C must be a multiple of 8");
  const int64_t total = a.npix * (a.C / 8);
  if (a.dtype == LEDB200_BF16)
    affine_relu_kernel<__nv_bfloat16><<<grid_for(total), kThreads, 0, st>>>(a);
  else
    affine_relu_kernel<float><<<grid_for(total), kThreads, 0, st>>>(a);
  LEDB_LAUNCH_OK("affine_relu_kernel");
  return LEDB200_OK;
}

int launch_nchw_to_nhwc(const float* in, void* out, int out_dtype, int N, int C, int H, int W, cudaStream_t st) {
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)ceil_div64(HW, 32), ceil_div(C, 32), N), block(32, 8);
  if (out_dtype == LEDB200_BF16)
    nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(in, (__nv_bfloat16*)out, C, HW);
  else
    nchw_to_nhwc_kernel<float><<<grid, block, 0, st>>>(in, (float*)out, C, HW);
  LEDB_LAUNCH_OK("nchw_to_nhwc_kernel");
  return LEDB200_OK;
}

int launch_nhwc_to_nchw(const void* in, int in_dtype, float* out, int N, int C, int H, int W, int in_ld, cudaStream_t st) {
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)ceil_div64(HW, 32), ceil_div(C, 32), N), block(32, 8);
  if (in_dtype == LEDB200_BF16)
    nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)in, out, C, HW, in_ld);
  else if (in_dtype == 5)   // IEEE fp16 (ladder rungs; debug fetch only)
    nhwc_to_nchw_kernel<__half><<<grid, block, 0, st>>>((const __half*)in, out, C, HW, in_ld);
  else
    nhwc_to_nchw_kernel<float><<<grid, block, 0, st>>>((const float*)in, out, C, HW, in_ld);
  LEDB_LAUNCH_OK("nhwc_to_nchw_kernel");
  return LEDB200_OK;
}

}  // namespace ledb
