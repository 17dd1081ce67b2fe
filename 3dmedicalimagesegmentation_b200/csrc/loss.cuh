// Loss kernels: softmax-Dice + cross-entropy (MONAI DiceCELoss(to_onehot_y=True, softmax=True), call site
// unetr_segmentation_3d.py:404,222) and the Bradley-Terry pairwise ranking loss
// (unetr_ranking_pretraining_3d.py:59-133 triplets, :202-217 loss).  HBM-bound: one pass over the logits
// for the forward, one read + one write for the backward.
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"

namespace b200 {

// ------------------------------------------------------------------ DiceCE
// acc layout (double): [B][C][3] = (I = sum p*t, P = sum p, G = sum t), then [B*C*3] = sum_v -(log p_target)
template <int CMAX>
__global__ void __launch_bounds__(256) dicece_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                         int C, long V, double* __restrict__ acc, int nBC) {
  int b = blockIdx.y;
  float aI[CMAX], aP[CMAX], aG[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) aI[c] = aP[c] = aG[c] = 0.f;
  float ce = 0.f;
  const float* lg = logits + (long)b * C * V;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
    float l[CMAX];
    float mx = -INFINITY, ly = 0.f;
    int y = (int)labels[(long)b * V + v];
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { l[c] = lg[(long)c * V + v]; mx = fmaxf(mx, l[c]); if (c == y) ly = l[c]; }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { l[c] = __expf(l[c] - mx); s += l[c]; }
    float inv = 1.f / s;
    ce += logf(s) - (ly - mx);
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        float p = l[c] * inv;
        aP[c] += p;
        if (c == y) { aI[c] += p; aG[c] += 1.f; }
      }
  }
  __shared__ float red[8][3 * CMAX + 1];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    float x0 = warp_sum(aI[c]), x1 = warp_sum(aP[c]), x2 = warp_sum(aG[c]);
    if (lane == 0) { red[w][3 * c] = x0; red[w][3 * c + 1] = x1; red[w][3 * c + 2] = x2; }
  }
  ce = warp_sum(ce);
  if (lane == 0) red[w][3 * CMAX] = ce;
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C + 1; i += blockDim.x) {
    int src = (i < 3 * C) ? i : 3 * CMAX;
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k][src];
    if (i < 3 * C) atomicAdd(acc + ((long)b * C) * 3 + i, t);
    else atomicAdd(acc + (long)nBC * 3, t);
  }
}
// out[0]=loss, out[1]=dice term, out[2]=ce term ; coef[b][c] = (a, bb) with d dice / d p = a*t + bb
static __global__ void dicece_finalize_kernel(const double* __restrict__ acc, int B, int C, long V, float* __restrict__ out,
                                       float* __restrict__ coef) {
  __shared__ double sd[256];
  double local = 0.0;
  int nBC = B * C;
  for (int i = threadIdx.x; i < nBC; i += blockDim.x) {
    double I = acc[3 * i], P = acc[3 * i + 1], G = acc[3 * i + 2];
    double den = G + P + 1e-5;
    local += 1.0 - (2.0 * I + 1e-5) / den;
    coef[2 * i] = (float)(-2.0 / (nBC * den));
    coef[2 * i + 1] = (float)((2.0 * I + 1e-5) / (nBC * den * den));
  }
  sd[threadIdx.x] = local;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) sd[threadIdx.x] += sd[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) {
    double dice = sd[0] / nBC, ce = acc[(long)nBC * 3] / ((double)B * V);
    out[0] = (float)(dice + ce); out[1] = (float)dice; out[2] = (float)ce;
  }
}
// dlogits_k = up * [ p_k (w_k - sum_c p_c w_c) + (p_k - t_k)/(B V) ],  w_c = a_c t_c + b_c
template <int CMAX>
__global__ void __launch_bounds__(256) dicece_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                         const float* __restrict__ coef, const float* __restrict__ upstream,
                                                         int B, int C, long V, float* __restrict__ dlogits) {
  int b = blockIdx.y;
  float up = upstream ? upstream[0] : 1.f;
  float invBV = 1.f / ((float)B * (float)V);
  __shared__ float sc[2 * CMAX];
  if (threadIdx.x < 2 * C) sc[threadIdx.x] = coef[(long)b * C * 2 + threadIdx.x];
  __syncthreads();
  const float* lg = logits + (long)b * C * V;
  float* dl = dlogits + (long)b * C * V;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
    float l[CMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { l[c] = lg[(long)c * V + v]; mx = fmaxf(mx, l[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { l[c] = __expf(l[c] - mx); s += l[c]; }
    float inv = 1.f / s;
    int y = (int)labels[(long)b * V + v];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { l[c] *= inv; dot += l[c] * (sc[2 * c] * (c == y ? 1.f : 0.f) + sc[2 * c + 1]); }
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        float t = (c == y) ? 1.f : 0.f;
        float wk = sc[2 * c] * t + sc[2 * c + 1];
        dl[(long)c * V + v] = up * (l[c] * (wk - dot) + (l[c] - t) * invBV);
      }
  }
}


// ---- 4 voxels per thread (16-byte loads per class plane; V % 4 == 0): 4x the bytes in flight of the scalar kernels above
template <int CMAX>
__global__ void __launch_bounds__(256) dicece_fwd4_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                          int C, long V, double* __restrict__ acc, int nBC) {
  int b = blockIdx.y;
  float aI[CMAX], aP[CMAX], aG[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) aI[c] = aP[c] = aG[c] = 0.f;
  float ce = 0.f;
  const float* lg = logits + (long)b * C * V;
  const long V4 = V >> 2;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < V4; q += (long)gridDim.x * blockDim.x) {
    float4 l[CMAX];
    const float4 yl = *reinterpret_cast<const float4*>(labels + (long)b * V + 4 * q);
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) l[c] = *reinterpret_cast<const float4*>(lg + (long)c * V + 4 * q);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int y = (int)(e == 0 ? yl.x : e == 1 ? yl.y : e == 2 ? yl.z : yl.w);
      float x[CMAX];
      float mx = -INFINITY, ly = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) { x[c] = e == 0 ? l[c].x : e == 1 ? l[c].y : e == 2 ? l[c].z : l[c].w; mx = fmaxf(mx, x[c]); if (c == y) ly = x[c]; }
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) { x[c] = __expf(x[c] - mx); s += x[c]; }
      const float inv = 1.f / s;
      ce += logf(s) - (ly - mx);
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          float p = x[c] * inv;
          aP[c] += p;
          if (c == y) { aI[c] += p; aG[c] += 1.f; }
        }
    }
  }
  __shared__ float red[8][3 * CMAX + 1];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    float x0 = warp_sum(aI[c]), x1 = warp_sum(aP[c]), x2 = warp_sum(aG[c]);
    if (lane == 0) { red[w][3 * c] = x0; red[w][3 * c + 1] = x1; red[w][3 * c + 2] = x2; }
  }
  ce = warp_sum(ce);
  if (lane == 0) red[w][3 * CMAX] = ce;
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C + 1; i += blockDim.x) {
    int src = (i < 3 * C) ? i : 3 * CMAX;
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k][src];
    if (i < 3 * C) atomicAdd(acc + ((long)b * C) * 3 + i, t);
    else atomicAdd(acc + (long)nBC * 3, t);
  }
}
template <int CMAX>
__global__ void __launch_bounds__(256) dicece_bwd4_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                          const float* __restrict__ coef, const float* __restrict__ upstream,
                                                          int B, int C, long V, float* __restrict__ dlogits) {
  int b = blockIdx.y;
  float up = upstream ? upstream[0] : 1.f;
  float invBV = 1.f / ((float)B * (float)V);
  __shared__ float sc[2 * CMAX];
  if (threadIdx.x < 2 * C) sc[threadIdx.x] = coef[(long)b * C * 2 + threadIdx.x];
  __syncthreads();
  const float* lg = logits + (long)b * C * V;
  float* dl = dlogits + (long)b * C * V;
  const long V4 = V >> 2;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < V4; q += (long)gridDim.x * blockDim.x) {
    float4 l[CMAX];
    const float4 yl = *reinterpret_cast<const float4*>(labels + (long)b * V + 4 * q);
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) l[c] = *reinterpret_cast<const float4*>(lg + (long)c * V + 4 * q);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int y = (int)(e == 0 ? yl.x : e == 1 ? yl.y : e == 2 ? yl.z : yl.w);
      float x[CMAX];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) { x[c] = e == 0 ? l[c].x : e == 1 ? l[c].y : e == 2 ? l[c].z : l[c].w; mx = fmaxf(mx, x[c]); }
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) { x[c] = __expf(x[c] - mx); s += x[c]; }
      const float inv = 1.f / s;
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) { x[c] *= inv; dot += x[c] * (sc[2 * c] * (c == y ? 1.f : 0.f) + sc[2 * c + 1]); }
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          float t = (c == y) ? 1.f : 0.f;
          float wk = sc[2 * c] * t + sc[2 * c + 1];
          float g = up * (x[c] * (wk - dot) + (x[c] - t) * invBV);
          if (e == 0) l[c].x = g; else if (e == 1) l[c].y = g; else if (e == 2) l[c].z = g; else l[c].w = g;
        }
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) *reinterpret_cast<float4*>(dl + (long)c * V + 4 * q) = l[c];
  }
}

// ------------------------------------------------------------------ DiceCE, staged through shared memory by bulk copies
// The kernels above read 14 class planes straight into registers: every thread has 14 dependent-latency loads per voxel and only
// ~6 voxels of work, and they reach 17 % (forward) / 23 % (backward) of the HBM roofline (round 1, ncu).  Here a CTA walks tiles of
// kDiceTile voxels: ONE thread issues C + 1 `cp.async.bulk` copies per tile (a class plane row or the label row, 2 KB each) into a
// 3-stage smem ring guarded by mbarriers, so 3 x 30 KB per CTA are in flight with no registers tied up, and the 256 threads compute
// two voxels each from conflict-free smem rows.  The backward overwrites the stage in place with dlogits and hands the rows back
// to the copy engine (`cp.async.bulk.global.shared::cta`), so its global stores are 2 KB bursts as well.
static constexpr int kDiceTile = 512, kDiceStages = 3;
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
static inline size_t dice_staged_smem(int C) { return (size_t)kDiceStages * (C + 1) * kDiceTile * sizeof(float) + 64; }

// tiles of one sample are dealt round-robin to the CTAs of that sample's grid row; V % kDiceTile == 0
template <int CMAX, bool BWD>
__global__ void __launch_bounds__(256) dicece_staged_kernel(const float* __restrict__ logits, const float* __restrict__ labels, int C, long V,
                                                            double* __restrict__ acc, int nBC, const float* __restrict__ coef,
                                                            const float* __restrict__ upstream, int B, float* __restrict__ dlogits) {
  extern __shared__ __align__(128) uint8_t dsm[];
  const int b = blockIdx.y;
  const uint32_t stage_floats = (uint32_t)(C + 1) * kDiceTile, stage_bytes = stage_floats * 4;
  float* ring = reinterpret_cast<float*>(dsm);
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + (size_t)kDiceStages * stage_bytes);
  __shared__ float sc[2 * CMAX];
  __shared__ float red[8][3 * CMAX + 1];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < kDiceStages; ++s) tc::mbar_init(full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (BWD && tid < 2 * C) sc[tid] = coef[(long)b * C * 2 + tid];
  __syncthreads();
  const long tiles = V / kDiceTile;
  const float* lg = logits + (long)b * C * V;
  const float* lb = labels + (long)b * V;
  float* dl = BWD ? dlogits + (long)b * C * V : nullptr;
  auto issue = [&](long tile, int s) {       // thread 0 only
    const uint32_t base = tc::smem_u32(ring + (size_t)s * stage_floats);
    tc::mbar_expect_tx(full + s, stage_bytes);
    for (int c = 0; c < C; ++c) bulk_g2s(base + (uint32_t)c * kDiceTile * 4, lg + (long)c * V + tile * kDiceTile, kDiceTile * 4, full + s);
    bulk_g2s(base + (uint32_t)C * kDiceTile * 4, lb + tile * kDiceTile, kDiceTile * 4, full + s);
  };
  if (tid == 0)
    for (int s = 0; s < kDiceStages; ++s) { const long t = blockIdx.x + (long)s * gridDim.x; if (t < tiles) issue(t, s); }
  float aI[CMAX], aP[CMAX], aG[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) aI[c] = aP[c] = aG[c] = 0.f;
  float ce = 0.f;
  const float up = BWD ? (upstream ? upstream[0] : 1.f) : 0.f;
  const float invBV = 1.f / ((float)B * (float)V);
  int it = 0;
  for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
    const int s = it % kDiceStages;
    tc::mbar_wait(full + s, (uint32_t)(it / kDiceStages) & 1u);
    float* st = ring + (size_t)s * stage_floats;
    {
      // two ADJACENT voxels per thread: 8-byte smem accesses, and two independent softmax chains to hide the MUFU / FMA latencies
      // (ncu on the one-voxel form: 492 instructions per voxel, stalls dominated by fixed-latency dependencies at 16 warps per SM)
      const int v = 2 * tid;
      float2 l[CMAX];
      float mx0 = -INFINITY, mx1 = -INFINITY, ly0 = 0.f, ly1 = 0.f;
      const float2 yy = *reinterpret_cast<const float2*>(st + C * kDiceTile + v);
      const int y0 = (int)yy.x, y1 = (int)yy.y;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          l[c] = *reinterpret_cast<const float2*>(st + c * kDiceTile + v);
          mx0 = fmaxf(mx0, l[c].x); mx1 = fmaxf(mx1, l[c].y);
          if (c == y0) ly0 = l[c].x;
          if (c == y1) ly1 = l[c].y;
        }
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) { l[c].x = __expf(l[c].x - mx0); l[c].y = __expf(l[c].y - mx1); s0 += l[c].x; s1 += l[c].y; }
      const float inv0 = 1.f / s0, inv1 = 1.f / s1;
      if (!BWD) {
        ce += (logf(s0) - (ly0 - mx0)) + (logf(s1) - (ly1 - mx1));
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) {
            const float p0 = l[c].x * inv0, p1 = l[c].y * inv1;
            aP[c] += p0 + p1;
            const float m0 = c == y0 ? 1.f : 0.f, m1 = c == y1 ? 1.f : 0.f;
            aI[c] = fmaf(m0, p0, fmaf(m1, p1, aI[c]));
            aG[c] += m0 + m1;
          }
      } else {
        float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) {
            l[c].x *= inv0; l[c].y *= inv1;
            dot0 = fmaf(l[c].x, sc[2 * c] * (c == y0 ? 1.f : 0.f) + sc[2 * c + 1], dot0);
            dot1 = fmaf(l[c].y, sc[2 * c] * (c == y1 ? 1.f : 0.f) + sc[2 * c + 1], dot1);
          }
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) {
            const float t0 = (c == y0) ? 1.f : 0.f, t1 = (c == y1) ? 1.f : 0.f;
            const float w0 = sc[2 * c] * t0 + sc[2 * c + 1], w1 = sc[2 * c] * t1 + sc[2 * c + 1];
            float2 o;
            o.x = up * (l[c].x * (w0 - dot0) + (l[c].x - t0) * invBV);
            o.y = up * (l[c].y * (w1 - dot1) + (l[c].y - t1) * invBV);
            *reinterpret_cast<float2*>(st + c * kDiceTile + v) = o;     // in place: only this thread touches columns v, v + 1
          }
      }
    }
    if (BWD) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();                                   // every thread is done with stage s
    if (tid == 0) {
      const long nxt = tile + (long)kDiceStages * gridDim.x;
      if (BWD) {
        const uint32_t base = tc::smem_u32(st);
        for (int c = 0; c < C; ++c) bulk_s2g(dl + (long)c * V + tile * kDiceTile, base + (uint32_t)c * kDiceTile * 4, kDiceTile * 4);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // stage s is refilled one iteration later (when its stores have read it): refill the PREVIOUS stage now
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        const long prev_nxt = nxt - gridDim.x;
        if (it >= 1 && prev_nxt < tiles) issue(prev_nxt, (it - 1) % kDiceStages);
      } else if (nxt < tiles) {
        issue(nxt, s);
      }
    }
  }
  if (BWD) {
    if (tid == 0) {
      // the last stage written has not been refilled (nothing left to load for it unless tiles remain: handled above for it-1 only)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
  const int lane = tid & 31, w = tid >> 5;
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    float x0 = warp_sum(aI[c]), x1 = warp_sum(aP[c]), x2 = warp_sum(aG[c]);
    if (lane == 0) { red[w][3 * c] = x0; red[w][3 * c + 1] = x1; red[w][3 * c + 2] = x2; }
  }
  ce = warp_sum(ce);
  if (lane == 0) red[w][3 * CMAX] = ce;
  __syncthreads();
  for (int i = tid; i < 3 * C + 1; i += blockDim.x) {
    const int src = (i < 3 * C) ? i : 3 * CMAX;
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k][src];
    if (i < 3 * C) atomicAdd(acc + ((long)b * C) * 3 + i, t);
    else atomicAdd(acc + (long)nBC * 3, t);
  }
}

// ------------------------------------------------------------------ DiceCE, multi-label variant
// MONAI DiceCELoss(to_onehot_y=False, sigmoid=True) as configured at unetr_segmentation_3d.py:480 for the 4-channel
// multi-hot BraTS target (seg:65-93): Dice on sigmoid probabilities against the float target [B][C][V], plus
// CrossEntropy(logits, argmax_c target) -- MONAI 0.6.0's `ce` takes the arg-max of a target that has as many channels as
// the prediction (first maximum wins, as torch.argmax does on CPU and CUDA).  Same accumulator layout and finalize kernel
// as the softmax variant; fwd reads logits + target once, bwd reads them once and writes dlogits once.
template <int CMAX>
__global__ void __launch_bounds__(256) dicece_sig_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                             int C, long V, double* __restrict__ acc, int nBC) {
  int b = blockIdx.y;
  float aI[CMAX], aP[CMAX], aG[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) aI[c] = aP[c] = aG[c] = 0.f;
  float ce = 0.f;
  const float* lg = logits + (long)b * C * V;
  const float* tg = target + (long)b * C * V;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
    float l[CMAX], t[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { l[c] = lg[(long)c * V + v]; t[c] = tg[(long)c * V + v]; }
    float mx = -INFINITY, tmax = -INFINITY, ly = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        mx = fmaxf(mx, l[c]);
        if (t[c] > tmax) { tmax = t[c]; ly = l[c]; }
        float p = 1.f / (1.f + __expf(-l[c]));
        aI[c] += p * t[c]; aP[c] += p; aG[c] += t[c];
      }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) s += __expf(l[c] - mx);
    ce += logf(s) - (ly - mx);
  }
  __shared__ float red[8][3 * CMAX + 1];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < CMAX; ++c) {
    float x0 = warp_sum(aI[c]), x1 = warp_sum(aP[c]), x2 = warp_sum(aG[c]);
    if (lane == 0) { red[w][3 * c] = x0; red[w][3 * c + 1] = x1; red[w][3 * c + 2] = x2; }
  }
  ce = warp_sum(ce);
  if (lane == 0) red[w][3 * CMAX] = ce;
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C + 1; i += blockDim.x) {
    int src = (i < 3 * C) ? i : 3 * CMAX;
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k][src];
    if (i < 3 * C) atomicAdd(acc + ((long)b * C) * 3 + i, t);
    else atomicAdd(acc + (long)nBC * 3, t);
  }
}
// dlogits_k = up * [ (a_k t_k + b_k) p_k (1 - p_k) + (softmax_k - [k == argmax t]) / (B V) ]
template <int CMAX>
__global__ void __launch_bounds__(256) dicece_sig_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                             const float* __restrict__ coef, const float* __restrict__ upstream,
                                                             int B, int C, long V, float* __restrict__ dlogits) {
  int b = blockIdx.y;
  float up = upstream ? upstream[0] : 1.f;
  float invBV = 1.f / ((float)B * (float)V);
  __shared__ float sc[2 * CMAX];
  if (threadIdx.x < 2 * C) sc[threadIdx.x] = coef[(long)b * C * 2 + threadIdx.x];
  __syncthreads();
  const float* lg = logits + (long)b * C * V;
  const float* tg = target + (long)b * C * V;
  float* dl = dlogits + (long)b * C * V;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long)gridDim.x * blockDim.x) {
    float l[CMAX], t[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { l[c] = lg[(long)c * V + v]; t[c] = tg[(long)c * V + v]; }
    float mx = -INFINITY, tmax = -INFINITY;
    int y = 0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { mx = fmaxf(mx, l[c]); if (t[c] > tmax) { tmax = t[c]; y = c; } }
    float e[CMAX];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { e[c] = __expf(l[c] - mx); s += e[c]; }
    float inv = 1.f / s;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        float p = 1.f / (1.f + __expf(-l[c]));
        float wk = sc[2 * c] * t[c] + sc[2 * c + 1];
        dl[(long)c * V + v] = up * (wk * p * (1.f - p) + (e[c] * inv - (c == y ? 1.f : 0.f)) * invBV);
      }
  }
}

// ------------------------------------------------------------------ Bradley-Terry ranking loss
// 16 slices: id = partition*4 + sample (samples ordered batch1[0], batch1[1], batch2[0], batch2[1], rank:80-84).
// A slice is [C, F0*F1] taken at index idx[partition] along the sliced spatial axis.
struct RankGeom {
  const float* src[4];  // sample base pointers (forward values)
  float* grad[4];       // gradient base pointers (same geometry) or null
  long sc, ss, sf0, sf1;  // element strides: channel, sliced axis, the two free axes
  int C, F0, F1;
  int idx[4];
  float temperature;
};
__device__ __forceinline__ long rank_off(const RankGeom& g, int part, int c, int f) {
  int f1 = f % g.F1, f0 = f / g.F1;
  return (long)c * g.sc + (long)g.idx[part] * g.ss + (long)f0 * g.sf0 + (long)f1 * g.sf1;
}
// gram[c][i][j] += sum_f s_i[c,f] s_j[c,f]      grid (C, splits), 256 threads = 16x16 pairs
static __global__ void __launch_bounds__(256) rank_gram_kernel(RankGeom g, double* __restrict__ gram) {
  __shared__ float tile[16][65];
  int c = blockIdx.x;
  int F = g.F0 * g.F1;
  int i = threadIdx.x >> 4, j = threadIdx.x & 15;
  float acc = 0.f;
  for (int f0 = blockIdx.y * 64; f0 < F; f0 += gridDim.y * 64) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      int s = e >> 6, ff = e & 63, f = f0 + ff;
      tile[s][ff] = (f < F) ? g.src[s & 3][rank_off(g, s >> 2, c, f)] : 0.f;
    }
    __syncthreads();
#pragma unroll 16
    for (int ff = 0; ff < 64; ++ff) acc = fmaf(tile[i][ff], tile[j][ff], acc);
    __syncthreads();
  }
  atomicAdd(gram + ((long)c * 16 + i) * 16 + j, (double)acc);
}
__device__ __forceinline__ void rank_triplet(int t, int& r, int& s, int& d) {
  int p = t / 144, rem = t % 144, pair = rem / 12, o = rem % 12;
  int rl = pair / 3, sl = pair % 3; if (sl >= rl) ++sl;
  int q = o >> 2; if (q >= p) ++q;
  r = p * 4 + rl; s = p * 4 + sl; d = q * 4 + (o & 3);
}
// per channel: cos matrix, loss contribution, and the 16x16 coefficient matrix coef[c][i][j] such that
// d loss / d s_i = sum_j coef[c][i][j] * s_j          grid C blocks of 576 threads... (use 576 = 18 warps)
static __global__ void __launch_bounds__(576) rank_loss_kernel(const double* __restrict__ gram, int C, float temperature,
                                                        double* __restrict__ loss, float* __restrict__ coef) {
  __shared__ float cosm[16][16], A[16][16], nrm[16], rawn[16];
  __shared__ float wsum[18];
  int c = blockIdx.x, t = threadIdx.x;
  const double* G = gram + (long)c * 256;
  if (t < 16) { float n = (float)sqrt(G[t * 16 + t]); rawn[t] = n; nrm[t] = fmaxf(n, 1e-6f); }
  if (t < 256) A[t >> 4][t & 15] = 0.f;
  __syncthreads();
  if (t < 256) cosm[t >> 4][t & 15] = (float)G[t] / (nrm[t >> 4] * nrm[t & 15]);
  __syncthreads();
  int r, s, d; rank_triplet(t, r, s, d);
  float z = (cosm[r][s] - cosm[r][d]) / temperature;
  float sp = (z > 0.f) ? log1pf(__expf(-z)) : (-z + log1pf(__expf(z)));  // softplus(-z), stable
  float dz = -1.f / (1.f + __expf(z)) / temperature / (float)C;       // d/dz softplus(-z) = -sigmoid(-z)
  atomicAdd(&A[r][s], dz);
  atomicAdd(&A[r][d], -dz);
  float ws = warp_sum(sp);
  if ((t & 31) == 0) wsum[t >> 5] = ws;
  __syncthreads();
  if (t == 0) { double tot = 0.0; for (int k = 0; k < 18; ++k) tot += wsum[k]; atomicAdd(loss, tot / C); }
  if (t < 256) {
    int i = t >> 4, j = t & 15;
    float out;
    if (i != j) {
      out = (A[i][j] + A[j][i]) / (nrm[i] * nrm[j]);
    } else {
      float acc = 0.f;
      for (int k = 0; k < 16; ++k) if (k != i) acc += (A[i][k] + A[k][i]) * cosm[i][k];
      out = (rawn[i] >= 1e-6f) ? -acc / (nrm[i] * nrm[i]) : 0.f;
    }
    coef[(long)c * 256 + t] = out;
  }
}
// grad slice i [c,f] = up * sum_j coef[c][i][j] s_j[c,f]         grid (C, ceil(F/256))
static __global__ void __launch_bounds__(256) rank_grad_kernel(RankGeom g, const float* __restrict__ coef,
                                                        const float* __restrict__ upstream) {
  __shared__ float cf[256];
  int c = blockIdx.x;
  cf[threadIdx.x] = coef[(long)c * 256 + threadIdx.x];
  __syncthreads();
  int F = g.F0 * g.F1;
  int f = blockIdx.y * 256 + threadIdx.x;
  if (f >= F) return;
  float up = upstream ? upstream[0] : 1.f;
  float sv[16];
#pragma unroll
  for (int s = 0; s < 16; ++s) sv[s] = g.src[s & 3][rank_off(g, s >> 2, c, f)];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) a = fmaf(cf[i * 16 + j], sv[j], a);
    g.grad[i & 3][rank_off(g, i >> 2, c, f)] = up * a;
  }
}

}  // namespace b200
