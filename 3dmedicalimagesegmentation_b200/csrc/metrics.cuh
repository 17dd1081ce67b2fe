// Validation tail of the segmentation script: MONAI DiceMetric / ConfusionMatrixMetric as configured at
// unetr_segmentation_3d.py:485-494 and used at :110-126, :153-188.  Everything the reference derives from the discretised
// prediction and the label -- Dice, precision, sensitivity per (sample, class) -- is a function of three integer counts per
// (sample, class): |y & p|, |p|, |y|.  One HBM pass produces them (from one-hot tensors as the reference passes them, or
// from the uint8 argmax mask + the integer label map, 5 B/voxel instead of 8*C), a one-block kernel turns them into the
// metrics with MONAI's NaN rules, and a second one-block kernel does the NaN-aware "mean" / "mean_batch" reductions.
#pragma once
#include "common.cuh"

namespace b200 {

// counts layout (double, integer-valued): [B][C][3] = (tp = sum y*p, np = sum p, ny = sum y)
// one-hot inputs [B][C][V] (any float values; MONAI passes 0/1): grid (chunks, B*C)
static __global__ void __launch_bounds__(256) seg_counts_onehot_kernel(const float* __restrict__ pred, const float* __restrict__ y,
                                                                       long V, double* __restrict__ counts) {
  const long base = (long)blockIdx.y * V;
  float tp = 0.f, np = 0.f, ny = 0.f;
  const bool vec = (V % 4 == 0) && ((((uintptr_t)pred | (uintptr_t)y) & 15) == 0);
  if (vec) {
    const float4* p4 = reinterpret_cast<const float4*>(pred + base);
    const float4* y4 = reinterpret_cast<const float4*>(y + base);
    for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < (V >> 2); q += (long)gridDim.x * blockDim.x) {
      float4 a = p4[q], b = y4[q];
      tp += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
      np += a.x + a.y + a.z + a.w;
      ny += b.x + b.y + b.z + b.w;
    }
  } else {
    for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
      float a = pred[base + v], b = y[base + v];
      tp += a * b; np += a; ny += b;
    }
  }
  __shared__ float red[8][3];
  tp = warp_sum(tp); np = warp_sum(np); ny = warp_sum(ny);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[w][0] = tp; red[w][1] = np; red[w][2] = ny; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(counts + (long)blockIdx.y * 3 + threadIdx.x, t);
  }
}

// label-map inputs: mask[B][V] uint8 (argmax class), labels[B][V] float holding integers; grid (chunks, B).
// Shared-memory integer histograms; values outside [0,C) count for no class (MONAI one_hot would raise).
static __global__ void __launch_bounds__(256) seg_counts_labels_kernel(const unsigned char* __restrict__ mask,
                                                                       const float* __restrict__ labels, int C, long V,
                                                                       double* __restrict__ counts) {
  __shared__ unsigned int h[3 * 32];
  for (int i = threadIdx.x; i < 3 * 32; i += blockDim.x) h[i] = 0u;
  __syncthreads();
  const long base = (long)blockIdx.y * V;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
    int p = mask[base + v], t = (int)labels[base + v];
    if (p < C) atomicAdd(&h[3 * p + 1], 1u);
    if ((unsigned)t < (unsigned)C) { atomicAdd(&h[3 * t + 2], 1u); if (t == p) atomicAdd(&h[3 * t], 1u); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x)
    if (h[i]) atomicAdd(counts + (long)blockIdx.y * C * 3 + i, (double)h[i]);
}

// dice[n][c] = 2 tp / (ny + np), NaN when ny == 0 (MONAI compute_meandice); confusion[n][c] = (tp, fp, tn, fn)
static __global__ void seg_metrics_kernel(const double* __restrict__ counts, int NC, double V, float* __restrict__ dice,
                                          float* __restrict__ confusion) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NC; i += gridDim.x * blockDim.x) {
    double tp = counts[3 * i], np = counts[3 * i + 1], ny = counts[3 * i + 2];
    if (dice) dice[i] = ny > 0.0 ? (float)(2.0 * tp / (ny + np)) : __int_as_float(0x7fc00000);
    if (confusion) {
      double fp = np - tp, fn = ny - tp;
      confusion[4 * i] = (float)tp; confusion[4 * i + 1] = (float)fp;
      confusion[4 * i + 2] = (float)(V - tp - fp - fn); confusion[4 * i + 3] = (float)fn;
    }
  }
}

// MONAI do_metric_reduction on f[N][C][K] (one block; N*C*K is tiny):
//   reduction 0 "mean":       out[K]    = NaN-aware mean over classes, then over the samples that had a non-NaN class
//   reduction 1 "mean_batch": out[C][K] = NaN-aware mean over samples
// not_nans receives the matching counts (same shape as out).
static __global__ void metric_reduce_kernel(const float* __restrict__ f, int N, int C, int K, int reduction, float* __restrict__ out,
                                            float* __restrict__ not_nans) {
  if (reduction == 0) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float sb = 0.f, nb = 0.f;
      for (int n = 0; n < N; ++n) {
        float s = 0.f, cnt = 0.f;
        for (int c = 0; c < C; ++c) { float v = f[((long)n * C + c) * K + k]; if (v == v) { s += v; cnt += 1.f; } }
        if (cnt > 0.f) { sb += s / cnt; nb += 1.f; }
      }
      out[k] = nb > 0.f ? sb / nb : 0.f;
      if (not_nans) not_nans[k] = nb;
    }
  } else {
    for (int i = threadIdx.x; i < C * K; i += blockDim.x) {
      float s = 0.f, cnt = 0.f;
      for (int n = 0; n < N; ++n) { float v = f[(long)n * C * K + i]; if (v == v) { s += v; cnt += 1.f; } }
      out[i] = cnt > 0.f ? s / cnt : 0.f;
      if (not_nans) not_nans[i] = cnt;
    }
  }
}

// MONAI compute_confusion_matrix_metric on rows of (tp, fp, tn, fn): metric 0 precision = tp/(tp+fp),
// 1 sensitivity (recall) = tp/(tp+fn); NaN where the denominator is 0.
static __global__ void confusion_metric_kernel(const float* __restrict__ cm, int rows, int metric, float* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
    float tp = cm[4 * i], fp = cm[4 * i + 1], fn = cm[4 * i + 3];
    float den = metric == 0 ? tp + fp : tp + fn;
    out[i] = den != 0.f ? tp / den : __int_as_float(0x7fc00000);
  }
}

}  // namespace b200
