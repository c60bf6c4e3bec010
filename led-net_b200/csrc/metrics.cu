// Confusion-matrix histogram (north_star kernel 5): warp-aggregated shared-memory atomics.
//
// Replaces IoUMetric.intersect_and_union (mmseg/evaluation/metrics/iou_metric.py:163-200: boolean
// mask gathers + 3x torch.histc in float32 + 3 D2H syncs per image) and
// calculate_confusion_matrix (tools/analysis_tools/confusion_matrix.py:66-74:
// bincount(K*gt + pred) after dropping gt == ignore_index; rows = GT, cols = prediction).
// The matrix is int64 [(K+1) x K]: row K collects GT values outside [0,K) that are not
// ignore_index - histc drops those from area_label but the pixel still counts in
// area_pred_label, and the spill row reproduces exactly that.  The four IoU histograms are
// diag / row sums / column sums of this matrix (host side).
//
// HBM-bound: 2 bytes per pixel in (uint8 pred + uint8 gt), nothing out.  Each thread reads 16
// pixels per 128-bit load; a warp aggregates equal bins with __match_any_sync so a blocky label
// map costs ~1 shared atomic per warp per load instead of 32.
#include "kernels.h"

namespace ledb {
namespace {

constexpr int kMaxBins = 256 * 255 + 256;   // (K+1)*K for K <= 255
constexpr int kCmThreads = 256;

__device__ __forceinline__ void warp_agg_add(unsigned int* hist, int bin, bool valid) {
  const unsigned active = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  const unsigned peers = __match_any_sync(active, bin);
  const int leader = __ffs(peers) - 1;
  if ((int)(threadIdx.x & 31) == leader) atomicAdd(&hist[bin], (unsigned)__popc(peers));
}

template <typename TP, typename TG>
__device__ __forceinline__ void ld_pix(const TP* pred, const TG* gt, int64_t i, int& p, int& g) {
  p = (int)pred[i];
  g = (int)gt[i];
}

// generic path: one pixel per thread per iteration
template <typename TP, typename TG>
__global__ void __launch_bounds__(kCmThreads)
confusion_kernel(const TP* __restrict__ pred, const TG* __restrict__ gt, int64_t n, int K, int ignore,
                 unsigned long long* __restrict__ cm) {
  extern __shared__ unsigned int hist[];
  const int bins = (K + 1) * K;
  for (int i = threadIdx.x; i < bins; i += kCmThreads) hist[i] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * kCmThreads;
  const int64_t n_round = ceil_div64(n, 32) * 32;   // keep warps converged for the ballots
  for (int64_t i = blockIdx.x * (int64_t)kCmThreads + threadIdx.x; i < n_round; i += stride) {
    bool valid = false;
    int bin = 0;
    if (i < n) {
      int p, g;
      ld_pix(pred, gt, i, p, g);
      if (g != ignore && p >= 0 && p < K) {
        valid = true;
        bin = ((g < 0 || g >= K) ? K : g) * K + p;
      }
    }
    warp_agg_add(hist, bin, valid);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += kCmThreads)
    if (hist[i]) atomicAdd(&cm[i], (unsigned long long)hist[i]);
}

// fast path: uint8 / uint8, 16 pixels per thread per iteration via 128-bit loads
__global__ void __launch_bounds__(kCmThreads)
confusion_u8x16_kernel(const uint4* __restrict__ pred, const uint4* __restrict__ gt, int64_t nvec, int K,
                       int ignore, unsigned long long* __restrict__ cm) {
  extern __shared__ unsigned int hist[];
  const int bins = (K + 1) * K;
  for (int i = threadIdx.x; i < bins; i += kCmThreads) hist[i] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * kCmThreads;
  const int64_t n_round = ceil_div64(nvec, 32) * 32;
  for (int64_t i = blockIdx.x * (int64_t)kCmThreads + threadIdx.x; i < n_round; i += stride) {
    uint4 pv = make_uint4(0, 0, 0, 0), gv = make_uint4(0, 0, 0, 0);
    const bool in = i < nvec;
    if (in) { pv = __ldg(pred + i); gv = __ldg(gt + i); }
    const unsigned pw[4] = {pv.x, pv.y, pv.z, pv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int p = (pw[w] >> (8 * b)) & 0xff, g = (gw[w] >> (8 * b)) & 0xff;
        const bool valid = in && g != ignore && p < K;
        const int bin = (g >= K ? K : g) * K + p;
        warp_agg_add(hist, bin, valid);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += kCmThreads)
    if (hist[i]) atomicAdd(&cm[i], (unsigned long long)hist[i]);
}

}  // namespace

int launch_confusion(const void* pred, const void* gt, int pred_dtype, int gt_dtype, int64_t n, int K,
                     int ignore_index, int64_t* cm, cudaStream_t st) {
  if (K < 1 || K > 255) return fail(LEDB200_EINVAL, "confusion: K must be in [1,255]");
  if (n < 0) return fail(LEDB200_EINVAL, "confusion: negative pixel count");
  if (n == 0) return LEDB200_OK;
  const int bins = (K + 1) * K;
  const size_t smem = bins * sizeof(unsigned int);
  if (bins > kMaxBins) return fail(LEDB200_EINVAL, "confusion: too many bins");
  auto* cmu = reinterpret_cast<unsigned long long*>(cm);
  // uint32 block-local counters: a block sees at most n/grid pixels; bound it below 2^32
  int grid = 148 * 8;
  if (pred_dtype == LEDB200_U8 && gt_dtype == LEDB200_U8 && ((uintptr_t)pred % 16 == 0) &&
      ((uintptr_t)gt % 16 == 0) && n >= 16) {
    const int64_t nvec = n / 16;
    if (smem > 48 * 1024)
      LEDB_CUDA_OK(cudaFuncSetAttribute(confusion_u8x16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int g = (int)std::min<int64_t>(grid, ceil_div64(nvec, kCmThreads));
    confusion_u8x16_kernel<<<g, kCmThreads, smem, st>>>((const uint4*)pred, (const uint4*)gt, nvec, K, ignore_index, cmu);
    LEDB_LAUNCH_OK("confusion_u8x16_kernel");
    const int64_t done = nvec * 16;
    if (done == n) return LEDB200_OK;
    pred = (const uint8_t*)pred + done;
    gt = (const uint8_t*)gt + done;
    n -= done;
  }
  int g = (int)std::min<int64_t>(grid, ceil_div64(n, kCmThreads));
#define LEDB_CM(TP, TG)                                                                                  \
  do {                                                                                                   \
    if (smem > 48 * 1024)                                                                                \
      LEDB_CUDA_OK(cudaFuncSetAttribute(confusion_kernel<TP, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    confusion_kernel<TP, TG><<<g, kCmThreads, smem, st>>>((const TP*)pred, (const TG*)gt, n, K, ignore_index, cmu); \
  } while (0)
  if (pred_dtype == LEDB200_U8 && gt_dtype == LEDB200_U8) LEDB_CM(uint8_t, uint8_t);
  else if (pred_dtype == LEDB200_U8 && gt_dtype == LEDB200_I64) LEDB_CM(uint8_t, int64_t);
  else if (pred_dtype == LEDB200_I64 && gt_dtype == LEDB200_U8) LEDB_CM(int64_t, uint8_t);
  else if (pred_dtype == LEDB200_I64 && gt_dtype == LEDB200_I64) LEDB_CM(int64_t, int64_t);
  else return fail(LEDB200_EINVAL, "confusion: dtypes must be U8 or I64");
#undef LEDB_CM
  LEDB_LAUNCH_OK("confusion_kernel");
  return LEDB200_OK;
}

}  // namespace ledb
