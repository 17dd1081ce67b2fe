// Sliding-window inference kernels (MONAI sliding_window_inference with constant blending; call sites
// unetr_segmentation_3d.py:109,143,694).  Window gather from the (virtually padded) volume, in-order
// overlap-add, and a normalise pass that divides by the analytic separable window count and crops
// the padding (optionally fusing the channel argmax).  fp32, NCDHW, bit-exact with the reference's
// accumulate-then-divide arithmetic.
#pragma once
#include "common.cuh"

namespace b200 {

struct SwGeom {
  int C;              // channels of the tensor being moved
  int D, H, W;        // un-padded volume size
  int pd, ph, pw;     // padding in front of each axis (symmetric pad of MONAI: half before)
  int PD, PH, PW;     // padded size
  int r0, r1, r2;     // roi
};
struct SwWindow { int b, s0, s1, s2; };
struct SwBatch { SwWindow w[16]; int n; };

// windows[k][c][roi] = vol[b][c][start+off - pad] or cval outside
static __global__ void sw_gather_kernel(const float* __restrict__ vol, float* __restrict__ windows, SwGeom g, SwBatch wb, float cval) {
  long per = (long)g.C * g.r0 * g.r1 * g.r2;
  long total = per * wb.n;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int k = (int)(e / per); long r = e % per;
    int z = (int)(r % g.r2); r /= g.r2; int y = (int)(r % g.r1); r /= g.r1; int x = (int)(r % g.r0); int c = (int)(r / g.r0);
    SwWindow w = wb.w[k];
    int d = w.s0 + x - g.pd, h = w.s1 + y - g.ph, ww = w.s2 + z - g.pw;
    float v = cval;
    if ((unsigned)d < (unsigned)g.D && (unsigned)h < (unsigned)g.H && (unsigned)ww < (unsigned)g.W)
      v = vol[((((long)w.b * g.C + c) * g.D + d) * g.H + h) * g.W + ww];
    windows[e] = v;
  }
}
// acc[b][c][start+off] += pred[k][c][off]   (one window per launch keeps the reference's summation order)
static __global__ void sw_accumulate_kernel(float* __restrict__ acc, const float* __restrict__ pred, SwGeom g, SwWindow w) {
  long per = (long)g.C * g.r0 * g.r1 * g.r2;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += (long)gridDim.x * blockDim.x) {
    long r = e;
    int z = (int)(r % g.r2); r /= g.r2; int y = (int)(r % g.r1); r /= g.r1; int x = (int)(r % g.r0); int c = (int)(r / g.r0);
    long o = ((((long)w.b * g.C + c) * g.PD + w.s0 + x) * g.PH + w.s1 + y) * g.PW + w.s2 + z;
    acc[o] += pred[e];
  }
}
// the same overlap-add for n windows of ONE batch item in one launch: a thread owns an accumulator voxel of the windows' bounding
// box and adds the predictions of the windows that cover it in window order -- the per-voxel sequence of float additions is
// exactly the one the per-window launches (and MONAI's loop) produce, but a voxel covered twice is read and written once.
struct SwBox { int x0, y0, z0, nx, ny, nz; };
// KMAX = unroll of the window loop (4 when the call has <= 4 windows: a quarter of the registers, twice the resident warps)
template <int VEC, int KMAX>   // VEC = 4: four consecutive z per lane (16-byte accesses; window z-starts, roi and the padded width are multiples of 4)
static __global__ void __launch_bounds__(256) sw_accumulate_multi_kernel(float* __restrict__ acc, const float* __restrict__ pred, SwGeom g, SwBatch wb, SwBox bx) {
  // grid = (row x of the bounding box, channel): a block owns one (c, x) plane, its warps the lines y = warp, warp + 8, ..., the lanes
  // walk z (coalesced accumulator accesses).  No division anywhere: window membership along x is decided once per block, along y
  // once per line, and only the z test is left per element (the first form, one warp per flat (c, x, y) index with 64-bit div/mod
  // per line, was instruction-bound: 0.45 ms per call of 4 windows for 450 MB).  All loads of a voxel are issued before the additions;
  // the additions happen in window order (bit-identical to the one-window-at-a-time loop).
  const int x = bx.x0 + blockIdx.x, c = blockIdx.y;
  const int b = wb.w[0].b, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int plane = g.r1 * g.r2;
  const long per = (long)g.r0 * plane;
  unsigned livex = 0;
  for (int k = 0; k < wb.n; ++k)
    if ((unsigned)(x - wb.w[k].s0) < (unsigned)g.r0) livex |= 1u << k;
  if (!livex) return;
  float* aplane = acc + (((long)b * g.C + c) * g.PD + x) * g.PH * g.PW;
  const float* pc = pred + (long)c * per;
  for (int yi = warp; yi < bx.ny; yi += nwarp) {
    const int y = bx.y0 + yi;
    unsigned live = 0;
    for (int k = 0; k < wb.n; ++k)
      if (((livex >> k) & 1u) && (unsigned)(y - wb.w[k].s1) < (unsigned)g.r1) live |= 1u << k;
    if (!live) continue;
    float* arow = aplane + (long)y * g.PW;
    for (int zi = lane * VEC; zi < bx.nz; zi += 32 * VEC) {
      const int z = bx.z0 + zi;
      float vals[KMAX][VEC]; unsigned has = 0;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < wb.n && ((live >> k) & 1u)) {
          const int dz = z - wb.w[k].s2;
          if ((unsigned)dz < (unsigned)g.r2) {
            const float* src = pc + (long)k * g.C * per + ((x - wb.w[k].s0) * plane + (y - wb.w[k].s1) * g.r2 + dz);
            if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(src); vals[k][0] = t.x; vals[k][1] = t.y; vals[k][VEC > 1 ? 2 : 0] = t.z; vals[k][VEC > 1 ? 3 : 0] = t.w; }
            else vals[k][0] = *src;
            has |= 1u << k;
          }
        }
      }
      if (has) {
        float a[VEC];
        if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(arow + z); a[0] = t.x; a[1] = t.y; a[VEC > 1 ? 2 : 0] = t.z; a[VEC > 1 ? 3 : 0] = t.w; }
        else a[0] = arow[z];
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if ((has >> k) & 1u) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) a[i] += vals[k][i];
          }
        if (VEC == 4) *reinterpret_cast<float4*>(arow + z) = make_float4(a[0], a[1], a[VEC > 1 ? 2 : 0], a[VEC > 1 ? 3 : 0]);
        else arow[z] = a[0];
      }
    }
  }
}
// ---- slab-owned accumulation (multi-GPU): a rank owns the padded rows [xoff, xoff + nrows) of ONE batch item and adds,
// in global window order, every contribution that touches them: its own windows and the row-clipped pieces of lower ranks' windows
// that arrived over NVLink.  A piece is a window prediction restricted to padded rows [x_lo, x_hi); its buffer is [C][nx][r1][r2]
// with buffer row 0 = padded row xbase (own windows: the full prediction, xbase = s0, nx = r0; received pieces: xbase = x_lo).
// Per voxel the float additions happen in piece order = window order: bit-identical to the single-GPU loop.
struct SwPiece { const float* pred; int s0, s1, s2, x_lo, x_hi, nx, xbase; };
struct SwPieces { SwPiece p[16]; int n; };
template <int VEC, int KMAX>
static __global__ void __launch_bounds__(256) sw_accumulate_slab_kernel(float* __restrict__ acc, SwGeom g, SwPieces ps, SwBox bx, int xoff, int nrows) {
  // same mapping as sw_accumulate_multi_kernel: grid = (row x of the bounding box, channel), warps over y, lanes over z, no division
  const int x = bx.x0 + blockIdx.x, c = blockIdx.y;
  if ((unsigned)(x - xoff) >= (unsigned)nrows) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int plane = g.r1 * g.r2;
  unsigned livex = 0;
  for (int k = 0; k < ps.n; ++k)
    if (x >= ps.p[k].x_lo && x < ps.p[k].x_hi) livex |= 1u << k;
  if (!livex) return;
  float* aplane = acc + ((long)c * nrows + (x - xoff)) * g.PH * g.PW;
  for (int yi = warp; yi < bx.ny; yi += nwarp) {
    const int y = bx.y0 + yi;
    unsigned live = 0;
    for (int k = 0; k < ps.n; ++k)
      if (((livex >> k) & 1u) && (unsigned)(y - ps.p[k].s1) < (unsigned)g.r1) live |= 1u << k;
    if (!live) continue;
    float* arow = aplane + (long)y * g.PW;
    for (int zi = lane * VEC; zi < bx.nz; zi += 32 * VEC) {
      const int z = bx.z0 + zi;
      // all loads of the voxel(s) first (independent, up to 16 in flight), then the additions in piece order: the order of the float
      // additions is what makes the result bit-identical to the sequential loop, the order of the loads is free
      float vals[KMAX][VEC]; unsigned has = 0;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < ps.n && ((live >> k) & 1u)) {
          const int dz = z - ps.p[k].s2;
          if ((unsigned)dz < (unsigned)g.r2) {
            const float* src = ps.p[k].pred + ((long)c * ps.p[k].nx + (x - ps.p[k].xbase)) * plane + ((y - ps.p[k].s1) * g.r2 + dz);
            if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(src); vals[k][0] = t.x; vals[k][1] = t.y; vals[k][VEC > 1 ? 2 : 0] = t.z; vals[k][VEC > 1 ? 3 : 0] = t.w; }
            else vals[k][0] = *src;
            has |= 1u << k;
          }
        }
      }
      if (has) {
        float a[VEC];
        if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(arow + z); a[0] = t.x; a[1] = t.y; a[VEC > 1 ? 2 : 0] = t.z; a[VEC > 1 ? 3 : 0] = t.w; }
        else a[0] = arow[z];
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if ((has >> k) & 1u) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) a[i] += vals[k][i];
          }
        if (VEC == 4) *reinterpret_cast<float4*>(arow + z) = make_float4(a[0], a[1], a[VEC > 1 ? 2 : 0], a[VEC > 1 ? 3 : 0]);
        else arow[z] = a[0];
      }
    }
  }
}

struct SwStarts { int n0, n1, n2; int s0[64], s1[64], s2[64]; };
// which rows a normalise pass works on: un-padded rows [d0, d0 + nd) of the volume, read from an accumulator that holds padded rows
// [acc_xoff, acc_xoff + acc_rows) and written to an output tensor that holds un-padded rows [out_d0, out_d0 + out_rows)
struct SwSlab { int d0, nd, acc_xoff, acc_rows, out_d0, out_rows; };
// out[b][c][d][h][w] = acc[b][c][d+pd][h+ph][w+pw] / count ; optional argmax over c -> mask[b][d][h][w] ;
// optional validation tail (seg:110-126): with `labels` [B][D][H][W] (integer-valued floats) the argmax is
// compared with the label in the same pass and counts[b][c][3] += (|y&p|, |p|, |y|) -- the inputs of DiceMetric /
// ConfusionMatrixMetric -- so whole-volume evaluation never materialises one-hot tensors.
static __global__ void sw_finalize_kernel(const float* __restrict__ acc, float* __restrict__ out, unsigned char* __restrict__ mask,
                                   SwGeom g, SwStarts st, const float* __restrict__ labels, double* __restrict__ counts, SwSlab sl) {
  const long vox = (long)g.D * g.H * g.W;
  const int b = blockIdx.y;                       // one sample per grid row: the block's histogram belongs to one sample
  __shared__ unsigned int hist[3 * 32];
  if (counts) {
    for (int i = threadIdx.x; i < 3 * 32; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
  }
  // a block walks (d, h) lines, its threads the w of a line: one division per line and block instead of two 64-bit div/mod per voxel;
  // the window counts along x and y are per line, only the z count is per voxel
  const long lines = (long)sl.nd * g.H;
  for (long line = blockIdx.x; line < lines; line += gridDim.x) {
    const int dd = (int)(line / g.H), h = (int)(line - (long)dd * g.H), d = sl.d0 + dd;
    const int x = d + g.pd, y = h + g.ph;
    int c0 = 0, c1 = 0;
    for (int i = 0; i < st.n0; ++i) c0 += (x >= st.s0[i] && x < st.s0[i] + g.r0);
    for (int i = 0; i < st.n1; ++i) c1 += (y >= st.s1[i] && y < st.s1[i] + g.r1);
    const float* aline = acc + ((((long)b * g.C) * sl.acc_rows + (x - sl.acc_xoff)) * g.PH + y) * g.PW + g.pw;
    const long cstride_a = (long)sl.acc_rows * g.PH * g.PW;
    float* oline = out ? out + (((long)b * g.C) * sl.out_rows + (d - sl.out_d0)) * g.H * g.W + (long)h * g.W : nullptr;
    const long cstride_o = (long)sl.out_rows * g.H * g.W;
    const long r_line = ((long)d * g.H + h) * g.W;
    for (int w = threadIdx.x; w < g.W; w += blockDim.x) {
      const int z = w + g.pw;
      int c2 = 0;
      for (int i = 0; i < st.n2; ++i) c2 += (z >= st.s2[i] && z < st.s2[i] + g.r2);
      const float cnt = (float)(c0 * c1 * c2);
      float best = -INFINITY; int arg = 0;
      for (int c = 0; c < g.C; ++c) {
        const float v = aline[c * cstride_a + w] / cnt;
        if (oline) oline[c * cstride_o + w] = v;
        if (v > best) { best = v; arg = c; }
      }
      const long r0 = r_line + w;
      if (mask) mask[(long)b * vox + r0] = (unsigned char)arg;
      if (counts) {
        int t = (int)labels[(long)b * vox + r0];
        atomicAdd(&hist[3 * arg + 1], 1u);
        if ((unsigned)t < (unsigned)g.C) { atomicAdd(&hist[3 * t + 2], 1u); if (t == arg) atomicAdd(&hist[3 * t], 1u); }
      }
    }
  }
  if (counts) {
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * g.C; i += blockDim.x) if (hist[i]) atomicAdd(counts + (long)b * g.C * 3 + i, (double)hist[i]);
  }
}

}  // namespace b200
