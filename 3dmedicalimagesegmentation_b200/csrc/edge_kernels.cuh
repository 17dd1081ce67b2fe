// Kernels for the two "edge" layers whose channel counts are too small for tensor cores:
//   * encoder1's convolutions on the raw input volume (C_in = 1..4; unetr.py:90-98): conv1 3x3x3 and the conv3 1x1x1
//     residual projection computed together, straight from the NCDHW fp32 input (no bf16 rounding of the image),
//     with the InstanceNorm sums fused; and their weight gradients;
//   * the segmentation head (UnetOutBlock 1x1x1 + bias, unetr.py:175) fused with the last InstanceNorm/LeakyReLU/residual
//     pass, so logits are formed from fp32 activations, and its backward.
// All HBM-bound; one thread per voxel, coalesced along W, 16-byte channel-row accesses.
#pragma once
#include "elementwise.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------- encoder1 forward
// c1[v,co] = sum_{ci,tap} x[ci, v+tap-1] W1[co][ci][tap] ; c3[v,co] = sum_ci x[ci,v] W3[co][ci]
// stats1/stats3: double [N][CO][2] (sum, sumsq), zeroed by the caller.
template <class T, int CO>
__global__ void __launch_bounds__(256) conv_in_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W1, const float* __restrict__ W3,
                                                          int Cin, int D, int H, int W, typename RawOf<T>::type* __restrict__ c1, typename RawOf<T>::type* __restrict__ c3,
                                                          double* __restrict__ stats1, double* __restrict__ stats3) {
  typedef typename RawOf<T>::type TR;   // raw conv outputs (see RawOf)
  extern __shared__ float sw[];   // W1 as [ci][tap][CO], then W3 as [ci][CO]
  __shared__ float red[8][4 * CO];
  const long V = (long)D * H * W;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < Cin * 27 * CO; i += 256) { int co = i % CO, r = i / CO, tap = r % 27, ci = r / 27; sw[i] = W1[((long)co * Cin + ci) * 27 + tap]; }
  for (int i = threadIdx.x; i < Cin * CO; i += 256) { int co = i % CO, ci = i / CO; sw[Cin * 27 * CO + i] = W3[(long)co * Cin + ci]; }
  __syncthreads();
  const float* w3 = sw + Cin * 27 * CO;
  float s1[CO], q1[CO], s3[CO], q3[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) s1[c] = q1[c] = s3[c] = q3[c] = 0.f;
  for (long v = (long)blockIdx.x * 256 + threadIdx.x; v < V; v += (long)gridDim.x * 256) {
    const int vi = (int)v; const int w = vi % W, t = vi / W, h = t % H, d = t / H;   // 32-bit: 64-bit div/mod is emulated (~100 instr each)
    float a1[CO], a3[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) a1[c] = a3[c] = 0.f;
    for (int ci = 0; ci < Cin; ++ci) {
      const float* xc = x + ((long)n * Cin + ci) * V;
      const float* wc = sw + ci * 27 * CO;
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
        int dd = d + kd - 1;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          int hh = h + kh - 1;
          bool ok = (unsigned)dd < (unsigned)D && (unsigned)hh < (unsigned)H;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            int ww = w + kw - 1;
            float xv = (ok && (unsigned)ww < (unsigned)W) ? xc[((long)dd * H + hh) * W + ww] : 0.f;
            const float* wt = wc + ((kd * 3 + kh) * 3 + kw) * CO;
#pragma unroll
            for (int c = 0; c < CO; ++c) a1[c] = fmaf(xv, wt[c], a1[c]);
            if (kd == 1 && kh == 1 && kw == 1) {
#pragma unroll
              for (int c = 0; c < CO; ++c) a3[c] = fmaf(xv, w3[ci * CO + c], a3[c]);
            }
          }
        }
      }
    }
    constexpr int VN = Vec16<TR>::N;
    TR* o1 = c1 + ((long)n * V + v) * CO; TR* o3 = c3 + ((long)n * V + v) * CO;
#pragma unroll
    for (int c0 = 0; c0 < CO; c0 += VN) {
      Vec16<TR> p, r;
#pragma unroll
      for (int i = 0; i < VN; ++i) { p.v[i] = a1[c0 + i]; r.v[i] = a3[c0 + i]; }
      p.store(o1 + c0); r.store(o3 + c0);
    }
#pragma unroll
    for (int c = 0; c < CO; ++c) { s1[c] += a1[c]; q1[c] += a1[c] * a1[c]; s3[c] += a3[c]; q3[c] += a3[c] * a3[c]; }
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    float a = warp_sum(s1[c]), b = warp_sum(q1[c]), e = warp_sum(s3[c]), f = warp_sum(q3[c]);
    if (lane == 0) { red[wp][c] = a; red[wp][CO + c] = b; red[wp][2 * CO + c] = e; red[wp][3 * CO + c] = f; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * CO; i += 256) {
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += red[k][i];
    int which = i / CO, c = i % CO;
    double* dst = (which < 2 ? stats1 : stats3) + ((long)n * CO + c) * 2 + (which & 1);
    atomicAdd(dst, tot);
  }
}


// Same contract as conv_in_fwd_kernel, restructured for occupancy and smem traffic: a thread owns TWO voxels adjacent in w and EIGHT
// output channels (CO/8 threads per voxel pair), so a weight vector read from smem serves two voxels, accumulators + statistics fit
// in ~100 registers (2 CTAs/SM instead of 1) and every voxel's channels leave as one 16-byte store per thread.  W must be even.
template <class T, int CO>
__global__ void __launch_bounds__(256, 2) conv_in_fwd2_kernel(const float* __restrict__ x, const float* __restrict__ W1, const float* __restrict__ W3,
                                                              int Cin, int D, int H, int W, typename RawOf<T>::type* __restrict__ c1,
                                                              typename RawOf<T>::type* __restrict__ c3, double* __restrict__ stats1, double* __restrict__ stats3) {
  typedef typename RawOf<T>::type TR;
  static_assert(Vec16<TR>::N == 8 || Vec16<TR>::N == 4, "8 channels per thread = one or two 16-byte stores");
  constexpr int NCG = CO / 8;                 // channel groups = threads per voxel pair
  extern __shared__ __align__(16) float sw2[];   // W1 as [ci][tap][CO], then W3 as [ci][CO]
  __shared__ float red2[8][4 * CO];
  const long V = (long)D * H * W;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < Cin * 27 * CO; i += 256) { int co = i % CO, r = i / CO, tap = r % 27, ci = r / 27; sw2[i] = W1[((long)co * Cin + ci) * 27 + tap]; }
  for (int i = threadIdx.x; i < Cin * CO; i += 256) { int co = i % CO, ci = i / CO; sw2[Cin * 27 * CO + i] = W3[(long)co * Cin + ci]; }
  __syncthreads();
  const float* w3 = sw2 + Cin * 27 * CO;
  const int cg = threadIdx.x % NCG, c0 = cg * 8;
  const int W2 = W >> 1;
  const long npairs = V >> 1;
  float s1[8], q1[8], s3[8], q3[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) s1[c] = q1[c] = s3[c] = q3[c] = 0.f;
  for (long pi = (long)blockIdx.x * (256 / NCG) + threadIdx.x / NCG; pi < npairs; pi += (long)gridDim.x * (256 / NCG)) {
    const int pv = (int)pi; const int w = (pv % W2) * 2, t = pv / W2, h = t % H, d = t / H;
    float a1[2][8], a3[2][8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a1[0][c] = a1[1][c] = a3[0][c] = a3[1][c] = 0.f;
    for (int ci = 0; ci < Cin; ++ci) {
      const float* xc = x + ((long)n * Cin + ci) * V;
      const float* wc = sw2 + ci * 27 * CO + c0;
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
        const int dd = d + kd - 1;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int hh = h + kh - 1;
          const bool ok = (unsigned)dd < (unsigned)D && (unsigned)hh < (unsigned)H;
          const float* xr = xc + ((long)dd * H + hh) * W + w;       // only dereferenced when ok
          float xv[4];
          xv[0] = (ok && w > 0) ? xr[-1] : 0.f;
          if (ok) { float2 m = *reinterpret_cast<const float2*>(xr); xv[1] = m.x; xv[2] = m.y; } else { xv[1] = xv[2] = 0.f; }
          xv[3] = (ok && w + 2 < W) ? xr[2] : 0.f;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float* wt = wc + ((kd * 3 + kh) * 3 + kw) * CO;
            const float4 wa = *reinterpret_cast<const float4*>(wt), wb = *reinterpret_cast<const float4*>(wt + 4);
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int c = 0; c < 8; ++c) { a1[0][c] = fmaf(xv[kw], wv[c], a1[0][c]); a1[1][c] = fmaf(xv[kw + 1], wv[c], a1[1][c]); }
          }
          if (kd == 1 && kh == 1) {
            const float4 wa = *reinterpret_cast<const float4*>(w3 + ci * CO + c0), wb = *reinterpret_cast<const float4*>(w3 + ci * CO + c0 + 4);
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int c = 0; c < 8; ++c) { a3[0][c] = fmaf(xv[1], wv[c], a3[0][c]); a3[1][c] = fmaf(xv[2], wv[c], a3[1][c]); }
          }
        }
      }
    }
    const long v = 2 * pi;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      TR* o1 = c1 + ((long)n * V + v + e) * CO + c0; TR* o3 = c3 + ((long)n * V + v + e) * CO + c0;
      constexpr int VN = Vec16<TR>::N;
#pragma unroll
      for (int q = 0; q < 8; q += VN) {
        Vec16<TR> p, r;
#pragma unroll
        for (int i = 0; i < VN; ++i) { p.v[i] = a1[e][q + i]; r.v[i] = a3[e][q + i]; }
        p.store(o1 + q); r.store(o3 + q);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) { s1[c] += a1[e][c]; q1[c] = fmaf(a1[e][c], a1[e][c], q1[c]); s3[c] += a3[e][c]; q3[c] = fmaf(a3[e][c], a3[e][c], q3[c]); }
    }
  }
  // reduce over the lanes that share a channel group (lane % NCG), then over the 8 warps
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float a = s1[c], b = q1[c], e = s3[c], f = q3[c];
    for (int o = NCG; o < 32; o <<= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o);
      e += __shfl_xor_sync(0xffffffffu, e, o); f += __shfl_xor_sync(0xffffffffu, f, o);
    }
    if (lane < NCG) { red2[wp][c0 + c] = a; red2[wp][CO + c0 + c] = b; red2[wp][2 * CO + c0 + c] = e; red2[wp][3 * CO + c0 + c] = f; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * CO; i += 256) {
    double tot = 0.0;
    for (int k = 0; k < 8; ++k) tot += red2[k][i];
    int which = i / CO, c = i % CO;
    double* dst = (which < 2 ? stats1 : stats3) + ((long)n * CO + c) * 2 + (which & 1);
    atomicAdd(dst, tot);
  }
}

// ---------------------------------------------------------------------------------------------- encoder1 weight gradients
// dW1[co][ci][tap] += sum_v dc1[v,co] x[ci,v+tap-1] ; dW3[co][ci] += sum_v dc3[v,co] x[ci,v]     (outputs zeroed by the caller)
// grid (voxel chunks, N, Cin); 8 warps, warp w owns taps {w, w+8, w+16, w+24<27}; warp 7's 4th slot does the 1x1 conv3.
template <class T, int CO>
__global__ void __launch_bounds__(256) conv_in_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dc1, const T* __restrict__ dc3,
                                                            int Cin, int D, int H, int W, long chunk, float* __restrict__ dW1, float* __restrict__ dW3) {
  const long V = (long)D * H * W;
  const int n = blockIdx.y, ci = blockIdx.z;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const float* xc = x + ((long)n * Cin + ci) * V;
  float acc[4][CO];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[s][c] = 0.f;
  const long v0 = (long)blockIdx.x * chunk, v1 = min(V, v0 + chunk);
  constexpr int VN = Vec16<T>::N;
  for (long v = v0 + lane; v < v1; v += 32) {
    const int vi = (int)v; const int w = vi % W, t = vi / W, h = t % H, d = t / H;   // 32-bit: 64-bit div/mod is emulated (~100 instr each)
    float g[CO], g3[CO];
    const T* r1 = dc1 + ((long)n * V + v) * CO;
#pragma unroll
    for (int c0 = 0; c0 < CO; c0 += VN) { Vec16<T> p; p.load(r1 + c0);
#pragma unroll
      for (int i = 0; i < VN; ++i) g[c0 + i] = p.v[i]; }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      int tap = wp + 8 * s;
      if (tap < 27) {
        int kw = tap % 3, kh = (tap / 3) % 3, kd = tap / 9;
        int dd = d + kd - 1, hh = h + kh - 1, ww = w + kw - 1;
        float xv = ((unsigned)dd < (unsigned)D && (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W) ? xc[((long)dd * H + hh) * W + ww] : 0.f;
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[s][c] = fmaf(xv, g[c], acc[s][c]);
      } else if (wp == 7 && s == 3) {   // tap slot 31 is free: use it for conv3 (1x1x1)
        const T* r3 = dc3 + ((long)n * V + v) * CO;
#pragma unroll
        for (int c0 = 0; c0 < CO; c0 += VN) { Vec16<T> p; p.load(r3 + c0);
#pragma unroll
          for (int i = 0; i < VN; ++i) g3[c0 + i] = p.v[i]; }
        float xv = xc[v];
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[s][c] = fmaf(xv, g3[c], acc[s][c]);
      }
    }
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    int tap = wp + 8 * s;
    bool is3 = (wp == 7 && s == 3);
    if (tap >= 27 && !is3) continue;
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      float tot = warp_sum(acc[s][c]);
      if (lane == 0) {
        if (is3) atomicAdd(dW3 + (long)c * Cin + ci, tot);
        else atomicAdd(dW1 + ((long)c * Cin + ci) * 27 + tap, tot);
      }
    }
  }
}


// Same contract as conv_in_wgrad_kernel, restructured like conv_in_fwd2: a thread owns TWO voxels adjacent in w, EIGHT output
// channels and the NINE taps of one kd plane (blockIdx.z = role = (channel octet, kd)), i.e. 72 (+8 for the 1x1x1 conv in the kd = 1
// role) register accumulators fed by 2 x 16-byte gradient loads and 12 input loads per iteration; one block reduction at the end.
template <class T, int CO>
__global__ void __launch_bounds__(256, 2) conv_in_wgrad2_kernel(const float* __restrict__ x, const T* __restrict__ dc1, const T* __restrict__ dc3,
                                                                int Cin, int D, int H, int W, long pairs_per_block, float* __restrict__ dW1,
                                                                float* __restrict__ dW3) {
  constexpr int NCG = CO / 8;
  __shared__ float red[8][80];
  const long V = (long)D * H * W;
  const int n = blockIdx.y;
  const int role = blockIdx.z % (3 * NCG), ci = blockIdx.z / (3 * NCG);
  const int cg = role % NCG, kd = role / NCG, c0 = cg * 8;
  const float* xc = x + ((long)n * Cin + ci) * V;
  const int W2 = W >> 1;
  const long npairs = V >> 1;
  const long p0 = (long)blockIdx.x * pairs_per_block, p1 = min(npairs, p0 + pairs_per_block);
  float acc[9][8], acc3[8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) acc3[c] = 0.f;
  constexpr int VN = Vec16<T>::N;
  for (long pi = p0 + threadIdx.x; pi < p1; pi += 256) {
    const int pv = (int)pi; const int w = (pv % W2) * 2, t = pv / W2, h = t % H, d = t / H;
    const long v = 2 * pi;
    float g[2][8];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const T* r1 = dc1 + ((long)n * V + v + e) * CO + c0;
#pragma unroll
      for (int q = 0; q < 8; q += VN) { Vec16<T> pk; pk.load(r1 + q);
#pragma unroll
        for (int i = 0; i < VN; ++i) g[e][q + i] = pk.v[i]; }
    }
    const int dd = d + kd - 1;
    const bool okd = (unsigned)dd < (unsigned)D;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hh = h + kh - 1;
      const bool ok = okd && (unsigned)hh < (unsigned)H;
      const float* xr = xc + ((long)dd * H + hh) * W + w;       // only dereferenced when ok
      float xv[4];
      xv[0] = (ok && w > 0) ? xr[-1] : 0.f;
      if (ok) { float2 m = *reinterpret_cast<const float2*>(xr); xv[1] = m.x; xv[2] = m.y; } else { xv[1] = xv[2] = 0.f; }
      xv[3] = (ok && w + 2 < W) ? xr[2] : 0.f;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[kh * 3 + kw][c] = fmaf(xv[kw], g[0][c], fmaf(xv[kw + 1], g[1][c], acc[kh * 3 + kw][c]));
      if (kd == 1 && kh == 1) {      // conv3 (1x1x1): centre voxel values xv[1], xv[2] with the dc3 gradients
        float g3[2][8];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const T* r3 = dc3 + ((long)n * V + v + e) * CO + c0;
#pragma unroll
          for (int q = 0; q < 8; q += VN) { Vec16<T> pk; pk.load(r3 + q);
#pragma unroll
            for (int i = 0; i < VN; ++i) g3[e][q + i] = pk.v[i]; }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) acc3[c] = fmaf(xv[1], g3[0][c], fmaf(xv[2], g3[1][c], acc3[c]));
      }
    }
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) { float s = warp_sum(acc[t][c]); if (lane == 0) red[wp][t * 8 + c] = s; }
#pragma unroll
  for (int c = 0; c < 8; ++c) { float s = warp_sum(acc3[c]); if (lane == 0) red[wp][72 + c] = s; }
  __syncthreads();
  for (int i = threadIdx.x; i < 80; i += 256) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += red[k][i];
    const int c = c0 + (i & 7);
    if (i < 72) atomicAdd(dW1 + ((long)c * Cin + ci) * 27 + kd * 9 + (i >> 3), tot);
    else if (kd == 1) atomicAdd(dW3 + (long)c * Cin + ci, tot);
  }
}

// ---------------------------------------------------------------------------------------------- fused last norm pass + head
// d0[v,c] = lrelu(norm(c2)[v,c] + norm(c3)[v,c]) (stored as T for the backward), logits[n][k][v] = sum_c d0_fp32[c] Wh[k][c] + bh[k]
template <class T, int CO>
__global__ void __launch_bounds__(256) in_apply_head_kernel(const typename RawOf<T>::type* __restrict__ c2, float* __restrict__ mr2, const typename RawOf<T>::type* __restrict__ c3,
                                                            float* __restrict__ mr3, T* __restrict__ d0, const float* __restrict__ Wh,
                                                            const float* __restrict__ bh, int ncls, long V, float* __restrict__ logits,
                                                            const double* __restrict__ acc2, const double* __restrict__ acc3, double invV) {
  __shared__ float sW[32 * CO], sb[32], sm[4 * CO];
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < ncls * CO; i += 256) sW[i] = Wh[i];
  if (threadIdx.x < ncls) sb[threadIdx.x] = bh[threadIdx.x];
  // (mean, rstd) of both inputs: from the (sum, sumsq) accumulators when given (in_moments; block 0 stores them for the backward)
  for (int i = threadIdx.x; i < 2 * CO; i += 256) {
    const int which = i / CO, c = i % CO;
    const double* acc = which ? acc3 : acc2; float* mr = which ? mr3 : mr2;
    float m, r;
    if (acc) { in_moments(acc + 2 * ((long)n * CO + c), invV, m, r); if (blockIdx.x == 0) { mr[((long)n * CO + c) * 2] = m; mr[((long)n * CO + c) * 2 + 1] = r; } }
    else { m = mr[((long)n * CO + c) * 2]; r = mr[((long)n * CO + c) * 2 + 1]; }
    sm[which * 2 * CO + 2 * c] = m; sm[which * 2 * CO + 2 * c + 1] = r;
  }
  __syncthreads();
  constexpr int VN = Vec16<T>::N;
  for (long v = (long)blockIdx.x * 256 + threadIdx.x; v < V; v += (long)gridDim.x * 256) {
    const long row = ((long)n * V + v) * CO;
    float o[CO];
#pragma unroll
    for (int c0 = 0; c0 < CO; c0 += VN) {
      Vec16<typename RawOf<T>::type> a, b; Vec16<T> r; a.load(c2 + row + c0); b.load(c3 + row + c0);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        int c = c0 + i;
        float val = lrelu((a.v[i] - sm[2 * c]) * sm[2 * c + 1] + (b.v[i] - sm[2 * CO + 2 * c]) * sm[2 * CO + 2 * c + 1]);
        o[c] = val; r.v[i] = val;
      }
      r.store(d0 + row + c0);
    }
    for (int k = 0; k < ncls; ++k) {
      float acc = sb[k];
#pragma unroll
      for (int c = 0; c < CO; ++c) acc = fmaf(o[c], sW[k * CO + c], acc);
      logits[((long)n * ncls + k) * V + v] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------- head backward
// g[v,c] = sum_k dlogits[n][k][v] Wh[k][c] (stored T, channels-last); dWh[k][c] += sum_v dlogits d0 ; dbh[k] += sum_v dlogits
// block = 256 threads over 256 voxels per iteration; (k,c) pairs are owned by threads for the outer-product sums.
template <class T, int CO>
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dlogits, const T* __restrict__ d0, const float* __restrict__ Wh,
                                                       int ncls, long V, long chunk, T* __restrict__ g, float* __restrict__ dWh, float* __restrict__ dbh) {
  extern __shared__ float hsm[];
  float* sW = hsm;                                            // [ncls][CO]
  float (*sdl)[257] = reinterpret_cast<float (*)[257]>(hsm + ncls * CO);          // [k][voxel]
  float (*sd0)[CO + 1] = reinterpret_cast<float (*)[CO + 1]>(hsm + ncls * CO + ncls * 257);  // [voxel][c]
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < ncls * CO; i += 256) sW[i] = Wh[i];
  __syncthreads();
  const int pairs = ncls * CO;
  float accw[4] = {0.f, 0.f, 0.f, 0.f};   // up to 1024 (k,c) pairs: four per thread
  float accb = 0.f;
  constexpr int VN = Vec16<T>::N;
  const long v0 = (long)blockIdx.x * chunk, v1 = min(V, v0 + chunk);
  for (long vb = v0; vb < v1; vb += 256) {
    long v = vb + threadIdx.x;
    bool ok = v < v1;
    float dl[32];
    for (int k = 0; k < ncls; ++k) { dl[k] = ok ? dlogits[((long)n * ncls + k) * V + v] : 0.f; sdl[k][threadIdx.x] = dl[k]; }
    if (ok) {
      const long row = ((long)n * V + v) * CO;
#pragma unroll
      for (int c0 = 0; c0 < CO; c0 += VN) {
        Vec16<T> a, o; a.load(d0 + row + c0);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          sd0[threadIdx.x][c0 + i] = a.v[i];
          float s = 0.f;
          for (int k = 0; k < ncls; ++k) s = fmaf(dl[k], sW[k * CO + c0 + i], s);
          o.v[i] = s;
        }
        o.store(g + row + c0);
      }
    } else {
#pragma unroll
      for (int c = 0; c < CO; ++c) sd0[threadIdx.x][c] = 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      int pi = threadIdx.x + 256 * s;
      if (pi < pairs) {
        int k = pi / CO, c = pi % CO;
        float a = 0.f;
#pragma unroll 8
        for (int j = 0; j < 256; ++j) a = fmaf(sdl[k][j], sd0[j][c], a);
        accw[s] += a;
      }
    }
    if (threadIdx.x < ncls) { float a = 0.f; for (int j = 0; j < 256; ++j) a += sdl[threadIdx.x][j]; accb += a; }
    __syncthreads();
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) { int pi = threadIdx.x + 256 * s; if (pi < pairs && dWh) atomicAdd(dWh + pi, accw[s]); }
  if (threadIdx.x < ncls && dbh) atomicAdd(dbh + threadIdx.x, accb);
}


// Same contract as head_bwd_kernel for ncls <= 16, with the outer-product sums register-tiled: thread (tile, group) owns a
// 2(k) x 8(c) tile of dWh and walks the voxels j = group, group + ngroup, ... of the staged block with one 8-byte and two
// 16-byte smem loads per 16 FMAs (the kernel above does 2 scalar loads per FMA and is smem-bound at ~310 us).
template <class T, int CO>
__global__ void __launch_bounds__(256) head_bwd2_kernel(const float* __restrict__ dlogits, const T* __restrict__ d0, const float* __restrict__ Wh,
                                                        int ncls, long V, long chunk, T* __restrict__ g, float* __restrict__ dWh, float* __restrict__ dbh) {
  extern __shared__ __align__(16) float hsm2[];
  float* sW = hsm2;                      // [16][CO], rows >= ncls are zero
  float* sdl = sW + 16 * CO;             // [256][16]  (k fastest)
  float* sd0 = sdl + 256 * 16;           // [256][CO]
  const int n = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < 16 * CO; i += 256) sW[i] = (i / CO) < ncls ? Wh[i] : 0.f;
  constexpr int NCT = CO / 8, NTILE = 8 * NCT, NGROUP = 256 / NTILE;   // 8 k-pairs x CO/8 channel octets
  const int tile = tid % NTILE, grp = tid / NTILE;
  const int tk = tile % 8, tc = tile / 8;
  float acc[2][8];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  float accb[16];   // per-thread partial sums of dlogits over this thread's voxels (bias gradient)
#pragma unroll
  for (int k = 0; k < 16; ++k) accb[k] = 0.f;
  constexpr int VN = Vec16<T>::N;
  const long v0 = (long)blockIdx.x * chunk, v1 = min(V, v0 + chunk);
  __syncthreads();
  for (long vb = v0; vb < v1; vb += 256) {
    const long v = vb + tid;
    const bool ok = v < v1;
    float dl[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { dl[k] = (ok && k < ncls) ? dlogits[((long)n * ncls + k) * V + v] : 0.f; accb[k] += dl[k]; }
#pragma unroll
    for (int k = 0; k < 16; k += 4) *reinterpret_cast<float4*>(sdl + tid * 16 + k) = make_float4(dl[k], dl[k + 1], dl[k + 2], dl[k + 3]);
    if (ok) {
      const long row = ((long)n * V + v) * CO;
#pragma unroll
      for (int c0 = 0; c0 < CO; c0 += VN) {
        Vec16<T> a, o; a.load(d0 + row + c0);
#pragma unroll
        for (int i = 0; i < VN; i += 4) *reinterpret_cast<float4*>(sd0 + tid * CO + c0 + i) = make_float4(a.v[i], a.v[i + 1], a.v[i + 2], a.v[i + 3]);
#pragma unroll
        for (int i = 0; i < VN; ++i) o.v[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (k < ncls) {
#pragma unroll
            for (int i = 0; i < VN; i += 4) {
              const float4 wv = *reinterpret_cast<const float4*>(sW + k * CO + c0 + i);   // broadcast
              o.v[i] = fmaf(dl[k], wv.x, o.v[i]); o.v[i + 1] = fmaf(dl[k], wv.y, o.v[i + 1]);
              o.v[i + 2] = fmaf(dl[k], wv.z, o.v[i + 2]); o.v[i + 3] = fmaf(dl[k], wv.w, o.v[i + 3]);
            }
          }
        }
        o.store(g + row + c0);
      }
    } else {
#pragma unroll
      for (int c = 0; c < CO; c += 4) *reinterpret_cast<float4*>(sd0 + tid * CO + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (grp < NGROUP) {
#pragma unroll 4
      for (int j = grp; j < 256; j += NGROUP) {
        const float2 a = *reinterpret_cast<const float2*>(sdl + j * 16 + 2 * tk);
        const float4 b0 = *reinterpret_cast<const float4*>(sd0 + j * CO + 8 * tc), b1 = *reinterpret_cast<const float4*>(sd0 + j * CO + 8 * tc + 4);
        acc[0][0] = fmaf(a.x, b0.x, acc[0][0]); acc[0][1] = fmaf(a.x, b0.y, acc[0][1]); acc[0][2] = fmaf(a.x, b0.z, acc[0][2]); acc[0][3] = fmaf(a.x, b0.w, acc[0][3]);
        acc[0][4] = fmaf(a.x, b1.x, acc[0][4]); acc[0][5] = fmaf(a.x, b1.y, acc[0][5]); acc[0][6] = fmaf(a.x, b1.z, acc[0][6]); acc[0][7] = fmaf(a.x, b1.w, acc[0][7]);
        acc[1][0] = fmaf(a.y, b0.x, acc[1][0]); acc[1][1] = fmaf(a.y, b0.y, acc[1][1]); acc[1][2] = fmaf(a.y, b0.z, acc[1][2]); acc[1][3] = fmaf(a.y, b0.w, acc[1][3]);
        acc[1][4] = fmaf(a.y, b1.x, acc[1][4]); acc[1][5] = fmaf(a.y, b1.y, acc[1][5]); acc[1][6] = fmaf(a.y, b1.z, acc[1][6]); acc[1][7] = fmaf(a.y, b1.w, acc[1][7]);
      }
    }
    __syncthreads();
  }
  // block-level sum over the voxel groups (smem, reusing the staging area), then one atomic per (k, c) per block
  if (grp < NGROUP) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) sdl[(grp * NTILE + tile) * 16 + a * 8 + b] = acc[a][b];
  }
  __syncthreads();
  if (dWh) {
    for (int e = tid; e < NTILE * 16; e += 256) {
      float tot = 0.f;
#pragma unroll 4
      for (int gq = 0; gq < NGROUP; ++gq) tot += sdl[gq * NTILE * 16 + e];
      const int tl = e / 16, ab = e % 16;
      const int k = 2 * (tl % 8) + ab / 8, cc = 8 * (tl / 8) + ab % 8;
      if (k < ncls) atomicAdd(dWh + k * CO + cc, tot);
    }
  }
  if (dbh) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) { float t = warp_sum(accb[k]); if ((tid & 31) == 0) sdl[(tid >> 5) * 16 + k] = t; }
    __syncthreads();
    if (tid < ncls) { float t = 0.f; for (int wv = 0; wv < 8; ++wv) t += sdl[wv * 16 + tid]; atomicAdd(dbh + tid, t); }
  }
}

}  // namespace b200

namespace b200 {
static inline size_t head_bwd2_smem(int CO) { return sizeof(float) * ((size_t)16 * CO + 256 * 16 + 256 * (size_t)CO); }
static inline size_t head_bwd_smem(int ncls, int CO) { return sizeof(float) * ((size_t)ncls * CO + (size_t)ncls * 257 + 256 * (size_t)(CO + 1)); }
}
