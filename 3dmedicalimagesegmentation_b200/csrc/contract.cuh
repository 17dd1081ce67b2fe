// Generic CUDA-core contraction  D[m,n] = sum_k A(m,k) * B(n,k)  with functor operand loaders and
// epilogues.  This is the fp32 "parity mode" engine (logits within 1e-4 of the fp32 oracle need true
// fp32 FMAs -- tcgen05 has no fp32-input MMA) and the bring-up path for any op whose
// tcgen05 kernel is not in place yet.  Every conv / transposed conv / patch-embedding / linear layer of
// UNETR, forward, dgrad and wgrad, is one instantiation of this template.
#pragma once
#include "common.cuh"

namespace b200 {

// ------------------------------------------------------------------ operand loaders
// A-loaders:  float operator()(int batch, int m, int k)      B-loaders: float operator()(int batch, int n, int k)
// kOuterFast: the outer index (m or n) is the memory-contiguous one (drives the tile-load thread mapping).

template <class T, bool OuterFast = false>
struct LdStrided {
  static constexpr bool kOuterFast = OuterFast;
  const T* p; long so, sk, sb0, sb1; int nb1;
  __device__ __forceinline__ float operator()(int b, int o, int k) const {
    return to_f(p[(long)(b / nb1) * sb0 + (long)(b % nb1) * sb1 + (long)o * so + (long)k * sk]);
  }
};

// Channels-last voxel gather for k^3 convolutions (k in {1,3}, stride 1, zero pad (k-1)/2).
// get(v, j): v = flat voxel over [N,D,H,W]; j = tap*C+c (TapMajor) or c*taps+tap.
// sign=+1 reads x[v + (tap-pad)] (forward / wgrad), sign=-1 reads x[v - (tap-pad)] (dgrad).
template <class T, bool TapMajor>
struct ConvGather {
  const T* x; int D, H, W, pitch, coff, C, ks, sign;
  __device__ __forceinline__ float get(int v, int j) const {
    int taps = ks * ks * ks;
    int tap, c;
    if (TapMajor) { tap = j / C; c = j - tap * C; } else { c = j / taps; tap = j - c * taps; }
    int pad = ks >> 1;
    int kw = tap % ks, kh = (tap / ks) % ks, kd = tap / (ks * ks);
    int w = v % W; int t = v / W; int h = t % H; t /= H; int d = t % D; int n = t / D;
    d += sign * (kd - pad); h += sign * (kh - pad); w += sign * (kw - pad);
    if ((unsigned)d >= (unsigned)D || (unsigned)h >= (unsigned)H || (unsigned)w >= (unsigned)W) return 0.f;
    return to_f(x[((((long)n * D + d) * H + h) * W + w) * pitch + coff + c]);
  }
};

// Gather for the transposed conv k2 s2: get(v, j) with v = flat INPUT voxel [N,D,H,W], j = co*8+tap,
// reads y (the 2x up-sampled tensor, channels-last) at (2d+a, 2h+b, 2w+c).
template <class T>
struct ConvTGather {
  const T* y; int D, H, W, pitch, coff;
  __device__ __forceinline__ float get(int v, int j) const {
    int co = j >> 3, tap = j & 7;
    int w = v % W; int t = v / W; int h = t % H; t /= H; int d = t % D; int n = t / D;
    long pos = (((long)n * 2 * D + 2 * d + (tap >> 2)) * 2 * H + 2 * h + ((tap >> 1) & 1)) * 2 * W + 2 * w + (tap & 1);
    return to_f(y[pos * pitch + coff + co]);
  }
};

// 16^3 patch rows straight from the NCDHW fp32 volume (MONAI PatchEmbeddingBlock).
// get(tok, k): perceptron k = ((p1*16+p2)*16+p3)*C + c ; conv k = c*4096 + (p1*16+p2)*16+p3.
struct PatchGather {
  const float* x; int C, S0, S1, S2, g0, g1, g2, conv_order;
  __device__ __forceinline__ float get(int tok, int k) const {
    int c, p;
    if (conv_order) { c = k >> 12; p = k & 4095; } else { c = k % C; p = k / C; }
    int p3 = p & 15, p2 = (p >> 4) & 15, p1 = p >> 8;
    int d = tok % g2; int t = tok / g2; int w = t % g1; t /= g1; int h = t % g0; int b = t / g0;
    return x[((((long)b * C + c) * S0 + h * 16 + p1) * S1 + w * 16 + p2) * S2 + d * 16 + p3];
  }
};

// NCDHW fp32 tensor read as [voxel, channel]
struct NcdhwGather {
  const float* p; int C; long V;
  __device__ __forceinline__ float get(int v, int c) const {
    long b = v / V; long vv = v - b * V;
    return p[(b * C + c) * V + vv];
  }
};

// adapters: gather functor G::get(row, col) used in either operand role
template <class G, bool OuterFast> struct RowIsOuter {   // op(b, o, k) = g(o, k)
  static constexpr bool kOuterFast = OuterFast; G g;
  __device__ __forceinline__ float operator()(int, int o, int k) const { return g.get(o, k); }
};
template <class G, bool OuterFast> struct RowIsK {       // op(b, o, k) = g(k, o)
  static constexpr bool kOuterFast = OuterFast; G g;
  __device__ __forceinline__ float operator()(int, int o, int k) const { return g.get(k, o); }
};

// PyTorch Conv3d weight [Co][Ci][taps] (fp32 master) as the B operand.
//   forward: n = co, k = tap*Ci+ci          dgrad: n = ci, k = tap*Co+co
struct ConvWeightB {
  static constexpr bool kOuterFast = false;
  const float* w; int Ci, Co, taps, dgrad;
  __device__ __forceinline__ float operator()(int, int n, int k) const {
    if (!dgrad) { int tap = k / Ci, ci = k - tap * Ci; return w[((long)n * Ci + ci) * taps + tap]; }
    int tap = k / Co, co = k - tap * Co; return w[((long)co * Ci + n) * taps + tap];
  }
};

// ------------------------------------------------------------------ epilogues:  void operator()(b, m, n, acc)
// ACT_GELU_SAVE (forward of a layer that will be differentiated): out = gelu(v) and `preact` receives gelu'(v) instead of v -- the erf is
// shared, and the backward epilogue (ACT_MUL_SAVED: v *= usrc) is one multiply instead of erf + exp per element
enum { ACT_NONE = 0, ACT_GELU = 1, ACT_GELU_BWD = 2, ACT_GELU_SAVE = 3, ACT_MUL_SAVED = 4 };
__device__ __forceinline__ void gelu_and_grad(float x, float& y, float& dy) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  y = x * cdf;
  dy = cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// 4 consecutive elements <-> floats (16-byte fp32 / 8-byte 16-bit accesses): the coalesced tcgen05 epilogue hands every lane 4 columns
__device__ __forceinline__ void ep_ld4(const float* p, float* o) { float4 t = *reinterpret_cast<const float4*>(p); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }
__device__ __forceinline__ void ep_ld4(const bf16* p, float* o) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ void ep_ld4(const __half* p, float* o) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ void ep_st4(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void ep_st4(bf16* p, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t; t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}
__device__ __forceinline__ void ep_st4(__half* p, const float* v) {
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  uint2 t; t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

template <class TO>
struct EpStore {
  TO* out; long ld, sb0, sb1; int nb1;
  const float* bias; const float* resid; long ldr; TO* preact; const TO* usrc; int act; int accumulate; float alpha;
  int splitk_nbat;   // > 0: split-K launch -- b carries split*nbat + batch.  split_stride > 0: each split stores its raw fp32 partial
                     // tile at out + split*split_stride (the CONSUMER kernel sums the partials and adds bias / residual);
                     // split_stride == 0: out (fp32, pre-zeroed) is accumulated atomically and split 0 adds bias / residual
  long split_stride;
  __device__ __forceinline__ void split_store(int b, int m, int n0, const float* acc, int nvalid) const {
    const int sp = b / splitk_nbat; b -= sp * splitk_nbat;
    long o = (long)(b / nb1) * sb0 + (long)(b % nb1) * sb1 + (long)m * ld + n0;
    if (split_stride) {
      float* dst = reinterpret_cast<float*>(out) + sp * split_stride + o;
      if (nvalid == 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j] * alpha, acc[j + 1] * alpha, acc[j + 2] * alpha, acc[j + 3] * alpha);
      } else {
        for (int j = 0; j < nvalid; ++j) dst[j] = acc[j] * alpha;
      }
      return;
    }
    for (int j = 0; j < nvalid; ++j) {
      float v = acc[j] * alpha;
      if (sp == 0) { if (bias) v += bias[n0 + j]; if (resid) v += resid[(long)m * ldr + n0 + j]; }
      atomicAdd(reinterpret_cast<float*>(out + o + j), v);
    }
  }
  __device__ __forceinline__ void operator()(int b, int m, int n, float acc) const {
    if (splitk_nbat) { split_store(b, m, n, &acc, 1); return; }
    long o = (long)(b / nb1) * sb0 + (long)(b % nb1) * sb1 + (long)m * ld + n;
    float v = acc * alpha;
    if (bias) v += bias[n];
    if (act == ACT_GELU_SAVE) { float y, d; gelu_and_grad(v, y, d); preact[o] = from_f<TO>(d); v = y; }
    else {
      if (preact) preact[o] = from_f<TO>(v);
      if (act == ACT_GELU) v = gelu_erf(v);
      else if (act == ACT_GELU_BWD) v *= gelu_erf_grad(to_f(usrc[o]));
      else if (act == ACT_MUL_SAVED) v *= to_f(usrc[o]);
    }
    if (resid) v += resid[(long)m * ldr + n];
    if (accumulate) v += to_f(out[o]);
    out[o] = from_f<TO>(v);
  }
  // 16 consecutive columns of row m (tcgen05 epilogue): 16-byte loads/stores when the segment is full and aligned
  __device__ __forceinline__ void seg16(int b, int m, int n0, const float* acc, int nvalid) const {
    if (splitk_nbat) { split_store(b, m, n0, acc, nvalid); return; }
    long o = (long)(b / nb1) * sb0 + (long)(b % nb1) * sb1 + (long)m * ld + n0;
    bool fast = nvalid == 16 && ((reinterpret_cast<uintptr_t>(out + o) & 15) == 0) && (!preact || (reinterpret_cast<uintptr_t>(preact + o) & 15) == 0) &&
                (!usrc || (reinterpret_cast<uintptr_t>(usrc + o) & 15) == 0) && (!resid || (reinterpret_cast<uintptr_t>(resid + (long)m * ldr + n0) & 15) == 0) &&
                (!bias || (reinterpret_cast<uintptr_t>(bias + n0) & 15) == 0);
    if (!fast) {
#pragma unroll 1
      for (int j = 0; j < nvalid; ++j) (*this)(b, m, n0 + j, acc[j]);
      return;
    }
    constexpr int VN = Vec16<TO>::N;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = acc[j] * alpha;
    if (bias) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) { float4 t = *reinterpret_cast<const float4*>(bias + n0 + j); v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w; }
    }
    if (act == ACT_GELU_SAVE) {
      float d[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { float y; gelu_and_grad(v[j], y, d[j]); v[j] = y; }
#pragma unroll
      for (int j = 0; j < 16; j += VN) { Vec16<TO> t; for (int i = 0; i < VN; ++i) t.v[i] = d[j + i]; t.store(preact + o + j); }
    } else if (preact) {
#pragma unroll
      for (int j = 0; j < 16; j += VN) { Vec16<TO> t; for (int i = 0; i < VN; ++i) t.v[i] = v[j + i]; t.store(preact + o + j); }
    }
    if (act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
    } else if (act == ACT_GELU_BWD) {
#pragma unroll
      for (int j = 0; j < 16; j += VN) { Vec16<TO> t; t.load(usrc + o + j); for (int i = 0; i < VN; ++i) v[j + i] *= gelu_erf_grad(t.v[i]); }
    } else if (act == ACT_MUL_SAVED) {
#pragma unroll
      for (int j = 0; j < 16; j += VN) { Vec16<TO> t; t.load(usrc + o + j); for (int i = 0; i < VN; ++i) v[j + i] *= t.v[i]; }
    }
    if (resid) {
      const float* r = resid + (long)m * ldr + n0;
#pragma unroll
      for (int j = 0; j < 16; j += 4) { float4 t = *reinterpret_cast<const float4*>(r + j); v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w; }
    }
    if (accumulate) {
#pragma unroll
      for (int j = 0; j < 16; j += VN) { Vec16<TO> t; t.load(out + o + j); for (int i = 0; i < VN; ++i) v[j + i] += t.v[i]; }
    }
#pragma unroll
    for (int j = 0; j < 16; j += VN) { Vec16<TO> t; for (int i = 0; i < VN; ++i) t.v[i] = v[j + i]; t.store(out + o + j); }
  }
  // TMA-store epilogue of tc::gemm (tc_gemm.cuh): the final values of 16 consecutive columns of row m -- and, for ACT_GELU_SAVE / preact,
  // the second output -- are RETURNED instead of stored; the kernel parks them in swizzled smem boxes and the copy engine writes them.
  // Plain dense outputs only: one batch, no accumulate (the host side of tc::gemm checks).  Split-K launches with partial tiles
  // (split_stride > 0) return the raw partial sums.  `live` = row m exists (m < M): rows past the edge are clipped by the tensor map.
  __device__ __forceinline__ void values16(bool live, int m, int n0, const float* acc, float* v, float* pv) const {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = acc[j] * alpha;
    if (splitk_nbat) return;
    constexpr int VN = Vec16<TO>::N;
    const long o = (long)m * ld + n0;
    if (bias) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) { float4 t = *reinterpret_cast<const float4*>(bias + n0 + j); v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w; }
    }
    if (act == ACT_GELU_SAVE) {
#pragma unroll
      for (int j = 0; j < 16; ++j) { float y; gelu_and_grad(v[j], y, pv[j]); v[j] = y; }
    } else if (preact) {
#pragma unroll
      for (int j = 0; j < 16; ++j) pv[j] = v[j];
    }
    if (act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
    } else if (act == ACT_GELU_BWD) {
      if (live) {
#pragma unroll
        for (int j = 0; j < 16; j += VN) { Vec16<TO> t; t.load(usrc + o + j); for (int i = 0; i < VN; ++i) v[j + i] *= gelu_erf_grad(t.v[i]); }
      }
    } else if (act == ACT_MUL_SAVED) {
      if (live) {
#pragma unroll
        for (int j = 0; j < 16; j += VN) { Vec16<TO> t; t.load(usrc + o + j); for (int i = 0; i < VN; ++i) v[j + i] *= t.v[i]; }
      }
    }
    if (resid && live) {
      const float* r = resid + (long)m * ldr + n0;
#pragma unroll
      for (int j = 0; j < 16; j += 4) { float4 t = *reinterpret_cast<const float4*>(r + j); v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w; }
    }
  }
  // is this launch a plain dense store the TMA epilogue can take?  (16-byte aligned operands, no batch offsets, no accumulate)
  bool tma_store_ok(int nbatch) const {
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (nbatch != 1 || accumulate || (splitk_nbat && !split_stride) || (splitk_nbat && splitk_nbat != 1)) return false;
    if (!al(out) || (ld * sizeof(TO)) % 16 || (preact && !al(preact)) || (usrc && !al(usrc)) || (bias && !al(bias))) return false;
    if (resid && (!al(resid) || (ldr * 4) % 16)) return false;
    if (splitk_nbat && (split_stride * 4) % 16) return false;
    return true;
  }
  // 4 consecutive columns of row m (coalesced tcgen05 epilogue: 8 lanes cover 32 consecutive columns of one row)
  __device__ __forceinline__ void seg4(int b, int m, int n0, const float* acc, int nvalid) const {
    constexpr uintptr_t AO = 4 * sizeof(TO) - 1;
    if (splitk_nbat) {
      if (split_stride && nvalid == 4) {
        const int sp = b / splitk_nbat; const int bb = b - sp * splitk_nbat;
        float* dst = reinterpret_cast<float*>(out) + sp * split_stride + (long)(bb / nb1) * sb0 + (long)(bb % nb1) * sb1 + (long)m * ld + n0;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          *reinterpret_cast<float4*>(dst) = make_float4(acc[0] * alpha, acc[1] * alpha, acc[2] * alpha, acc[3] * alpha);
          return;
        }
      }
      split_store(b, m, n0, acc, nvalid);
      return;
    }
    long o = (long)(b / nb1) * sb0 + (long)(b % nb1) * sb1 + (long)m * ld + n0;
    bool fast = nvalid == 4 && ((reinterpret_cast<uintptr_t>(out + o) & AO) == 0) && (!preact || (reinterpret_cast<uintptr_t>(preact + o) & AO) == 0) &&
                (!usrc || (reinterpret_cast<uintptr_t>(usrc + o) & AO) == 0) && (!resid || (reinterpret_cast<uintptr_t>(resid + (long)m * ldr + n0) & 15) == 0) &&
                (!bias || (reinterpret_cast<uintptr_t>(bias + n0) & 15) == 0);
    if (!fast) {
#pragma unroll 1
      for (int j = 0; j < nvalid; ++j) (*this)(b, m, n0 + j, acc[j]);
      return;
    }
    float v[4] = {acc[0] * alpha, acc[1] * alpha, acc[2] * alpha, acc[3] * alpha};
    if (bias) { float4 t = *reinterpret_cast<const float4*>(bias + n0); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
    if (act == ACT_GELU_SAVE) {
      float d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { float y; gelu_and_grad(v[j], y, d[j]); v[j] = y; }
      ep_st4(preact + o, d);
    } else if (preact) ep_st4(preact + o, v);
    if (act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
    } else if (act == ACT_GELU_BWD) {
      float u[4]; ep_ld4(usrc + o, u);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= gelu_erf_grad(u[j]);
    } else if (act == ACT_MUL_SAVED) {
      float u[4]; ep_ld4(usrc + o, u);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= u[j];
    }
    if (resid) { float4 t = *reinterpret_cast<const float4*>(resid + (long)m * ldr + n0); v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w; }
    if (accumulate) { float u[4]; ep_ld4(out + o, u); v[0] += u[0]; v[1] += u[1]; v[2] += u[2]; v[3] += u[3]; }
    ep_st4(out + o, v);
  }
  // Per-tile context of the coalesced epilogue: the batch offset (two integer divisions) and every alignment test are done once per
  // output tile instead of once per 4 columns (the epilogue warps are one per SM sub-partition: issue-bound).
  struct Tile { long ob; int fast; int sp; };
  __device__ __forceinline__ Tile tile(int b) const {
    constexpr uintptr_t AO = 4 * sizeof(TO) - 1;
    Tile t; t.sp = 0;
    int bb = b;
    if (splitk_nbat) { t.sp = b / splitk_nbat; bb = b - t.sp * splitk_nbat; }
    t.ob = (long)(bb / nb1) * sb0 + (long)(bb % nb1) * sb1;
    bool f;
    if (splitk_nbat) {
      f = split_stride != 0 && ((reinterpret_cast<uintptr_t>(reinterpret_cast<float*>(out) + t.sp * split_stride + t.ob) & 15) == 0) && (ld & 3) == 0;
    } else {
      f = ((reinterpret_cast<uintptr_t>(out + t.ob) & AO) == 0) && (ld & 3) == 0;
      if (preact) f = f && ((reinterpret_cast<uintptr_t>(preact + t.ob) & AO) == 0);
      if (usrc) f = f && ((reinterpret_cast<uintptr_t>(usrc + t.ob) & AO) == 0);
      if (resid) f = f && ((reinterpret_cast<uintptr_t>(resid) & 15) == 0) && (ldr & 3) == 0;
      if (bias) f = f && ((reinterpret_cast<uintptr_t>(bias) & 15) == 0);
    }
    t.fast = f;
    return t;
  }
  __device__ __forceinline__ void seg4t(const Tile& t, int b, int m, int n0, const float* acc, int nvalid) const {
    if (!t.fast || nvalid != 4 || (n0 & 3)) { seg4(b, m, n0, acc, nvalid); return; }
    const long o = t.ob + (long)m * ld + n0;
    if (splitk_nbat) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + t.sp * split_stride + o) = make_float4(acc[0] * alpha, acc[1] * alpha, acc[2] * alpha, acc[3] * alpha);
      return;
    }
    float v[4] = {acc[0] * alpha, acc[1] * alpha, acc[2] * alpha, acc[3] * alpha};
    if (bias) { float4 q = *reinterpret_cast<const float4*>(bias + n0); v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w; }
    if (act == ACT_GELU_SAVE) {
      float d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { float y; gelu_and_grad(v[j], y, d[j]); v[j] = y; }
      ep_st4(preact + o, d);
    } else if (preact) ep_st4(preact + o, v);
    if (act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
    } else if (act == ACT_GELU_BWD) {
      float u[4]; ep_ld4(usrc + o, u);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= gelu_erf_grad(u[j]);
    } else if (act == ACT_MUL_SAVED) {
      float u[4]; ep_ld4(usrc + o, u);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= u[j];
    }
    if (resid) { float4 q = *reinterpret_cast<const float4*>(resid + (long)m * ldr + n0); v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w; }
    if (accumulate) { float u[4]; ep_ld4(out + o, u); v[0] += u[0]; v[1] += u[1]; v[2] += u[2]; v[3] += u[3]; }
    ep_st4(out + o, v);
  }
};
template <class TO> static inline EpStore<TO> ep_plain(TO* out, long ld) {
  EpStore<TO> e; memset(&e, 0, sizeof(e)); e.out = out; e.ld = ld; e.nb1 = 1; e.alpha = 1.f; return e;
}

// Row-wise attention epilogues for tc::gemm (see row_op_of in tc_gemm.cuh).  Rows are [batch pair][m][ldw] of TP.
template <class TP>
struct EpSoftmaxRow {
  static constexpr int kRowOp = 1;
  TP* P; long sb0, sb1; int nb1; int ldw; float scale;
  __device__ __forceinline__ void store16(int b, int m, int c0, const float* v, int nvalid) const {
    TP* dst = P + (long)(b / nb1) * sb0 + (long)(b % nb1) * sb1 + (long)m * ldw + c0;
    constexpr int VN = Vec16<TP>::N;
    if (nvalid == 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
      for (int q = 0; q < 16; q += VN) { Vec16<TP> t; for (int i = 0; i < VN; ++i) t.v[i] = v[q + i]; t.store(dst + q); }
    } else {
      for (int j = 0; j < nvalid; ++j) dst[j] = from_f<TP>(v[j]);
    }
  }
  __device__ __forceinline__ void operator()(int, int, int, float) const {}
};
template <class TP>
struct EpSoftmaxBwdRow {
  static constexpr int kRowOp = 2;
  const TP* Pin; TP* dS; long sb0, sb1; int nb1; int ldw; float scale;
  __device__ __forceinline__ void load_p16(int b, int m, int c0, float* pr, int nvalid) const {
    const TP* src = Pin + (long)(b / nb1) * sb0 + (long)(b % nb1) * sb1 + (long)m * ldw + c0;
    constexpr int VN = Vec16<TP>::N;
    if (nvalid == 16 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll
      for (int q = 0; q < 16; q += VN) { Vec16<TP> t; t.load(src + q); for (int i = 0; i < VN; ++i) pr[q + i] = t.v[i]; }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) pr[j] = j < nvalid ? to_f(src[j]) : 0.f;
    }
  }
  __device__ __forceinline__ void store16(int b, int m, int c0, const float* v, int nvalid) const {
    TP* dst = dS + (long)(b / nb1) * sb0 + (long)(b % nb1) * sb1 + (long)m * ldw + c0;
    constexpr int VN = Vec16<TP>::N;
    if (nvalid == 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
      for (int q = 0; q < 16; q += VN) { Vec16<TP> t; for (int i = 0; i < VN; ++i) t.v[i] = v[q + i]; t.store(dst + q); }
    } else {
      for (int j = 0; j < nvalid; ++j) dst[j] = from_f<TP>(v[j]);
    }
  }
  __device__ __forceinline__ void operator()(int, int, int, float) const {}
};

struct EpPatch {  // tokens = acc + bias + position embedding
  float* out; int H, L; const float* bias; const float* pos;
  __device__ __forceinline__ void operator()(int, int m, int n, float acc) const {
    out[(long)m * H + n] = acc + bias[n] + pos[(long)(m % L) * H + n];
  }
  __device__ __forceinline__ void seg16(int, int m, int n0, const float* acc, int nvalid) const {
    const float* pr = pos + (long)(m % L) * H + n0; float* o = out + (long)m * H + n0;
    if (nvalid == 16 && (H & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 bb = *reinterpret_cast<const float4*>(bias + n0 + j), pp = *reinterpret_cast<const float4*>(pr + j);
        *reinterpret_cast<float4*>(o + j) = make_float4(acc[j] + bb.x + pp.x, acc[j + 1] + bb.y + pp.y, acc[j + 2] + bb.z + pp.z, acc[j + 3] + bb.w + pp.w);
      }
    } else {
      for (int j = 0; j < nvalid; ++j) o[j] = acc[j] + bias[n0 + j] + pr[j];
    }
  }
};

struct EpAtomic {  // split-K weight gradients
  float* out; long ld;
  __device__ __forceinline__ void operator()(int, int m, int n, float acc) const { atomicAdd(out + (long)m * ld + n, acc); }
};

struct EpAtomicTapRemap {  // GEMM column n = tap*Co+co -> PyTorch ConvTranspose3d weight column co*8+tap, atomic (split-K)
  float* out; int Co;
  __device__ __forceinline__ void operator()(int, int m, int n, float acc) const {
    int tap = n / Co, co = n - tap * Co;
    atomicAdd(out + (long)m * Co * 8 + co * 8 + tap, acc);
  }
};

template <class TO>
struct EpConvTScatter {  // m = input voxel, n = co*8+tap -> channels-last 2x up-sampled output
  TO* out; int D, H, W, pitch, coff;
  __device__ __forceinline__ void operator()(int, int m, int n, float acc) const {
    int co = n >> 3, tap = n & 7;
    int w = m % W; int t = m / W; int h = t % H; t /= H; int d = t % D; int nb = t / D;
    long pos = (((long)nb * 2 * D + 2 * d + (tap >> 2)) * 2 * H + 2 * h + ((tap >> 1) & 1)) * 2 * W + 2 * w + (tap & 1);
    out[pos * pitch + coff + co] = from_f<TO>(acc);
  }
};

template <class TO>
struct EpConvTScatterTap {  // m = input voxel, n = tap*Co + co (tap-major packed weight, Co % 16 == 0): 16 channels per 32-byte store
  TO* out; int D, H, W, pitch, coff, Co;
  __device__ __forceinline__ long pos_of(int m, int tap) const {
    int w = m % W; int t = m / W; int h = t % H; t /= H; int d = t % D; int nb = t / D;
    return (((long)nb * 2 * D + 2 * d + (tap >> 2)) * 2 * H + 2 * h + ((tap >> 1) & 1)) * 2 * W + 2 * w + (tap & 1);
  }
  __device__ __forceinline__ void operator()(int, int m, int n, float acc) const {
    int tap = n / Co, co = n - tap * Co;
    out[pos_of(m, tap) * pitch + coff + co] = from_f<TO>(acc);
  }
  __device__ __forceinline__ void seg16(int, int m, int n0, const float* acc, int nvalid) const {
    int tap = n0 / Co, co = n0 - tap * Co;
    TO* dst = out + pos_of(m, tap) * pitch + coff + co;
    constexpr int VN = Vec16<TO>::N;
#pragma unroll
    for (int q = 0; q < 16 / VN; ++q) {
      Vec16<TO> o;
#pragma unroll
      for (int j = 0; j < VN; ++j) o.v[j] = acc[q * VN + j];
      o.store(dst + q * VN);
    }
  }
};

struct EpHeadNcdhw {  // logits[b][co][v] = acc + bias[co]
  float* out; int Co; long V; const float* bias;
  __device__ __forceinline__ void operator()(int, int m, int n, float acc) const {
    long b = m / V; long vv = m - b * V;
    out[(b * Co + n) * V + vv] = acc + bias[n];
  }
};

// ------------------------------------------------------------------ the kernel
template <class AL, class BL, class EP, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) contract_kernel(AL al, BL bl, EP ep, int M, int N, int K, int nsplit, int kchunk) {
  constexpr int BK = 16;
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int batch = blockIdx.z / nsplit, split = blockIdx.z % nsplit;
  const int k_begin = split * kchunk, k_end = min(K, k_begin + kchunk);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tn = tid % (BN / TN), tm = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int e = tid; e < BM * BK; e += 256) {
      int ml, kl;
      if (AL::kOuterFast) { ml = e % BM; kl = e / BM; } else { kl = e % BK; ml = e / BK; }
      int m = m0 + ml, k = k0 + kl;
      As[kl][ml] = (m < M && k < k_end) ? al(batch, m, k) : 0.f;
    }
#pragma unroll
    for (int e = tid; e < BN * BK; e += 256) {
      int nl, kl;
      if (BL::kOuterFast) { nl = e % BN; kl = e / BN; } else { kl = e % BK; nl = e / BK; }
      int n = n0 + nl, k = k0 + kl;
      Bs[kl][nl] = (n < N && k < k_end) ? bl(batch, n, k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][tm * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tn * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + tm * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tn * TN + j;
      if (n < N) ep(batch, m, n, acc[i][j]);
    }
  }
}

// Host launcher.  nsplit>1 requires an atomic epilogue.
template <class AL, class BL, class EP>
static int launch_contract(const AL& al, const BL& bl, const EP& ep, int M, int N, int K, int batch, int nsplit,
                           cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  int kchunk = ((cdiv(K, nsplit) + 15) / 16) * 16;
  nsplit = cdiv(K, kchunk);
  if (N <= 16 && M <= 16) {
    dim3 g(1, 1, batch * nsplit);
    contract_kernel<AL, BL, EP, 16, 16, 1, 1><<<g, 256, 0, st>>>(al, bl, ep, M, N, K, nsplit, kchunk);
  } else if (N <= 16) {
    dim3 g(cdiv(M, 256), cdiv(N, 16), batch * nsplit);
    contract_kernel<AL, BL, EP, 256, 16, 4, 4><<<g, 256, 0, st>>>(al, bl, ep, M, N, K, nsplit, kchunk);
  } else if (N <= 32) {
    dim3 g(cdiv(M, 128), cdiv(N, 32), batch * nsplit);
    contract_kernel<AL, BL, EP, 128, 32, 4, 4><<<g, 256, 0, st>>>(al, bl, ep, M, N, K, nsplit, kchunk);
  } else {
    dim3 g(cdiv(M, 64), cdiv(N, 64), batch * nsplit);
    contract_kernel<AL, BL, EP, 64, 64, 4, 4><<<g, 256, 0, st>>>(al, bl, ep, M, N, K, nsplit, kchunk);
  }
  B200_LAUNCH_CHECK();
  return 0;
}

}  // namespace b200
