// SESP block as ONE kernel (north_star kernel 2: depthwise and low-channel convs on CUDA cores;
// SURVEY section 8a row B5).
//
// Replaces SESP.forward (mmseg/models/nn_layers/eesp.py:76-118) with its helper layers
// CBR / BR / CB / CDilated (mmseg/models/nn_layers/espnet_utils.py:8-145), eval mode, stride 1:
//   o1   = PReLU(BN(grouped 1x1 conv, groups = k))                      (proj_1x1)           n = nOut/k channels
//   b_i  = depthwise 3x3, dilation d_i, on o1;  S_i = b_i + S_{i-1}     (spp_dw + HFF)       i = 0..k-1
//   v_i  = depthwise 3x3, dilation d_i + 1, on S_i                      (spp_dw_v2, SESPV2)
//   e    = BN(grouped 1x1 conv(PReLU(BN(cat(v_0..v_{k-1})))))           (br_after_cat, conv_1x1_exp)
//   out  = PReLU(e + x)                                                 (residual when shapes match)
// The reference runs 12+ kernels and a torch.cat copy; here the input tile (with halo) is read once and
// the output written once - every intermediate lives in shared memory.  Group g of the expand conv
// consumes exactly branch g's n channels, so branches are independent after the HFF running sum.
//
// One CTA = one TS x TS output tile of one image, all channels; the depthwise stages walk the n
// channels in chunks of CH.  Zero padding is applied where the reference applies it: to o1 (first
// depthwise conv) and to S_i (second), i.e. both are literal zeros outside the image.
#include "kernels.h"

namespace ledb {
namespace {

constexpr int SESP_K = 4;
constexpr int SESP_THREADS = 256;

struct SespArgs {
  const void* in;
  void* out;
  const float* p;      // packed parameters, see sesp_param_offsets
  int N, H, W, nIn, nOut, n, cin_g, cpg;
  int d[SESP_K], d2[SESP_K];
  int v2, residual;
  int TS, R1, R2, CH;
  // offsets (floats) into p
  int o_wproj, o_pscale, o_pshift, o_pslope, o_wdw, o_wdw2, o_bscale, o_bshift, o_bslope, o_wexp, o_escale,
      o_eshift, o_aslope;
};

__device__ __forceinline__ float prelu(float v, float s) { return v > 0.f ? v : v * s; }

template <typename T>
__global__ void __launch_bounds__(SESP_THREADS) sesp_kernel(SespArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int TS = a.TS, R1 = a.R1, R2 = a.R2, CH = a.CH, n = a.n;
  const int E1 = TS + 2 * (R1 + R2), E2 = TS + 2 * R2;
  float* s1 = sm;                                  // [E1*E1][CH]
  float* S = s1 + E1 * E1 * CH;                    // [E2*E2][CH]
  float* V = S + E2 * E2 * CH;                     // [K][TS*TS][n]
  const int t = threadIdx.x;
  const int tiles_x = (a.W + TS - 1) / TS;
  const int ty0 = (blockIdx.x / tiles_x) * TS, tx0 = (blockIdx.x % tiles_x) * TS;
  const int img = blockIdx.y;
  const T* in = reinterpret_cast<const T*>(a.in) + (int64_t)img * a.H * a.W * a.nIn;
  T* out = reinterpret_cast<T*>(a.out) + (int64_t)img * a.H * a.W * a.nOut;
  const float* P = a.p;

  for (int c0 = 0; c0 < n; c0 += CH) {
    // ---- A: o1 on the (TS + 2(R1+R2))^2 halo region for channels c0..c0+CH
    for (int idx = t; idx < E1 * E1 * CH; idx += SESP_THREADS) {
      const int c = idx % CH, pix = idx / CH;
      const int gy = ty0 - (R1 + R2) + pix / E1, gx = tx0 - (R1 + R2) + pix % E1;
      float v = 0.f;
      if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
        const int ch = c0 + c;
        const int g = ch / a.cpg;
        const T* src = in + ((int64_t)gy * a.W + gx) * a.nIn + g * a.cin_g;
        const float* w = P + a.o_wproj + ch * a.cin_g;
        float acc = 0.f;
        for (int j = 0; j < a.cin_g; j += 8) {
          float xv[8];
          load8(src + j, xv);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc = fmaf(xv[q], __ldg(w + j + q), acc);
        }
        v = prelu(fmaf(acc, P[a.o_pscale + ch], P[a.o_pshift + ch]), P[a.o_pslope + ch]);
      }
      s1[idx] = v;
    }
    __syncthreads();
    for (int br = 0; br < SESP_K; ++br) {
      // ---- B1: S = (br ? S : 0) + depthwise(o1, dilation d[br]) on the (TS + 2 R2)^2 region
      const int d = a.d[br];
      for (int idx = t; idx < E2 * E2 * CH; idx += SESP_THREADS) {
        const int c = idx % CH, pix = idx / CH;
        const int py = pix / E2, px = pix % E2;
        const int gy = ty0 - R2 + py, gx = tx0 - R2 + px;
        float v = 0.f;
        if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
          const float* w = P + a.o_wdw + (br * n + c0 + c) * 9;
          const float* base = s1 + ((py + R1) * E1 + (px + R1)) * CH + c;
          float acc = br ? S[idx] : 0.f;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
              acc = fmaf(__ldg(w + kh * 3 + kw), base[((kh - 1) * d * E1 + (kw - 1) * d) * CH], acc);
          v = acc;
        }
        S[idx] = v;
      }
      __syncthreads();
      // ---- B2: v = depthwise(S, dilation d2[br]) (or S itself), then BN + PReLU of br_after_cat -> V
      const int d2 = a.d2[br];
      for (int idx = t; idx < TS * TS * CH; idx += SESP_THREADS) {
        const int c = idx % CH, pix = idx / CH;
        const int py = pix / TS, px = pix % TS;
        const float* base = S + ((py + R2) * E2 + (px + R2)) * CH + c;
        float v;
        if (a.v2) {
          const float* w = P + a.o_wdw2 + (br * n + c0 + c) * 9;
          v = 0.f;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
              v = fmaf(__ldg(w + kh * 3 + kw), base[((kh - 1) * d2 * E2 + (kw - 1) * d2) * CH], v);
        } else {
          v = base[0];
        }
        const int co = br * n + c0 + c;
        V[(br * TS * TS + pix) * n + c0 + c] = prelu(fmaf(v, P[a.o_bscale + co], P[a.o_bshift + co]), P[a.o_bslope + co]);
      }
      __syncthreads();
    }
  }
  // ---- C: grouped 1x1 expand + BN + residual + PReLU, 8 output channels per thread-iteration
  const int ogs = a.nOut / 8;
  for (int idx = t; idx < TS * TS * ogs; idx += SESP_THREADS) {
    const int og = idx % ogs, pix = idx / ogs;
    const int gy = ty0 + pix / TS, gx = tx0 + pix % TS;
    if (gy >= a.H || gx >= a.W) continue;
    const int co0 = og * 8;
    const int g = co0 / n;
    const float* v = V + (g * TS * TS + pix) * n;
    const float* w = P + a.o_wexp + co0 * n;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < n; j += 4) {
      const float4 x4 = *reinterpret_cast<const float4*>(v + j);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + q * n + j));
        acc[q] = fmaf(x4.x, w4.x, fmaf(x4.y, w4.y, fmaf(x4.z, w4.z, fmaf(x4.w, w4.w, acc[q]))));
      }
    }
    const int64_t pixoff = (int64_t)gy * a.W + gx;
    float r[8];
    if (a.residual) load8(in + pixoff * a.nIn + co0, r);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float e = fmaf(acc[q], P[a.o_escale + co0 + q], P[a.o_eshift + co0 + q]);
      if (a.residual) e += r[q];
      acc[q] = prelu(e, P[a.o_aslope + co0 + q]);
    }
    store8(out + pixoff * a.nOut + co0, acc);
  }
}

void sesp_offsets(SespArgs& a) {
  int o = 0;
  a.o_wproj = o; o += a.n * a.cin_g;
  a.o_pscale = o; o += a.n;
  a.o_pshift = o; o += a.n;
  a.o_pslope = o; o += a.n;
  a.o_wdw = o; o += SESP_K * a.n * 9;
  a.o_wdw2 = o; o += SESP_K * a.n * 9;
  a.o_bscale = o; o += a.nOut;
  a.o_bshift = o; o += a.nOut;
  a.o_bslope = o; o += a.nOut;
  a.o_wexp = o; o += a.nOut * a.n;
  a.o_escale = o; o += a.nOut;
  a.o_eshift = o; o += a.nOut;
  a.o_aslope = o;
}

}  // namespace
}  // namespace ledb

using namespace ledb;

extern "C" {

// number of floats of the packed parameter block for SESP(nIn, nOut) (k = 4):
//   w_proj[n][nIn/4], proj scale/shift/slope [n] x3, w_dw[4][n][9], w_dw2[4][n][9] (zeros when !v2),
//   br scale/shift/slope [nOut] x3, w_exp[nOut][n], exp scale/shift [nOut] x2, act slope [nOut]
int64_t ledb200_sesp_param_floats(int32_t nIn, int32_t nOut) {
  const int64_t n = nOut / SESP_K;
  return n * (nIn / SESP_K) + 3 * n + 2 * SESP_K * n * 9 + 3 * (int64_t)nOut + (int64_t)nOut * n + 3 * (int64_t)nOut;
}

int ledb200_sesp_forward(const void* in, void* out, int32_t dtype, int32_t N, int32_t H, int32_t W, int32_t nIn,
                         int32_t nOut, const int32_t* dilations4, int32_t v2, const float* params, void* stream) {
  if (!in || !out || !params || !dilations4) return fail(LEDB200_EINVAL, "sesp: null buffer");
  if (dtype != LEDB200_F32 && dtype != LEDB200_BF16) return fail(LEDB200_EINVAL, "sesp: dtype must be F32 or BF16");
  if (N < 1 || H < 1 || W < 1) return fail(LEDB200_EINVAL, "sesp: empty input");
  if (nOut % (SESP_K * 8) || nIn % (SESP_K * 8))
    return fail(LEDB200_EINVAL, "sesp: nIn and nOut must be multiples of 32 (k = 4 branches, 8-channel vectors)");
  SespArgs a;
  a.in = in; a.out = out; a.p = params; a.N = N; a.H = H; a.W = W; a.nIn = nIn; a.nOut = nOut;
  a.n = nOut / SESP_K; a.cin_g = nIn / SESP_K; a.cpg = a.n / SESP_K;
  a.v2 = v2 ? 1 : 0; a.residual = (nIn == nOut) ? 1 : 0;
  a.R1 = 0; a.R2 = 0;
  for (int i = 0; i < SESP_K; ++i) {
    if (dilations4[i] < 1 || dilations4[i] > 12) return fail(LEDB200_EINVAL, "sesp: dilation out of range [1,12]");
    a.d[i] = dilations4[i]; a.d2[i] = dilations4[i] + 1;
    a.R1 = a.R1 > a.d[i] ? a.R1 : a.d[i];
    if (a.v2) a.R2 = a.R2 > a.d2[i] ? a.R2 : a.d2[i];
  }
  sesp_offsets(a);
  // tile / chunk selection under the 227 KB shared-memory limit
  a.CH = a.n < 16 ? a.n : 16;
  size_t smem = 0;
  for (int ts : {16, 8, 4}) {
    a.TS = ts;
    const int E1 = ts + 2 * (a.R1 + a.R2), E2 = ts + 2 * a.R2;
    smem = sizeof(float) * ((size_t)E1 * E1 * a.CH + (size_t)E2 * E2 * a.CH + (size_t)SESP_K * ts * ts * a.n);
    if (smem <= 160 * 1024) break;
  }
  if (smem > 227 * 1024) return fail(LEDB200_EINVAL, "sesp: configuration does not fit shared memory");
  dim3 grid(ceil_div(W, a.TS) * ceil_div(H, a.TS), N);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LEDB200_BF16) {
    LEDB_CUDA_OK(cudaFuncSetAttribute(sesp_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sesp_kernel<__nv_bfloat16><<<grid, SESP_THREADS, smem, st>>>(a);
  } else {
    LEDB_CUDA_OK(cudaFuncSetAttribute(sesp_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sesp_kernel<float><<<grid, SESP_THREADS, smem, st>>>(a);
  }
  LEDB_LAUNCH_OK("sesp_kernel");
  return LEDB200_OK;
}

}  // extern "C"
