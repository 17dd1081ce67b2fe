// Fused multi-tensor AdamW (the reference calls torch.optim.AdamW at unetr_segmentation_3d.py:522,225-226 and
// unetr_ranking_pretraining_3d.py:466,214-215).  One launch updates every parameter that has a gradient:
//     p <- p*(1 - lr*wd);  m <- b1*m + (1-b1)*g;  v <- b2*v + (1-b2)*g^2;  p <- p - (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// (decoupled weight decay, no amsgrad: the arithmetic of torch.optim.AdamW).  Parameters whose grad is None are simply absent from
// the table, so their state does not move (H7).  HBM-bound: 28 bytes per element (read g,p,m,v; write p,m,v).
//
// Mirrors: the bf16 engine reads packed bf16 copies of the GEMM / conv weights (Exec::layout_packed).  A tensor with a non-null `s0`
// has its updated value written there as bf16 by THIS kernel (+2 bytes per element on 28), so a training step has no separate cast
// launch (round 1: multi_cast = 0.10 ms per step re-reading what AdamW had just written).  Only the plain-cast mirrors (ViT and
// patch-embedding weights, untransposed transposed-conv weights: 93 % of the parameters) are written here; the RE-LAID-OUT conv
// copies are scattered 2-byte stores when driven from the source order (measured: +0.25 ms inside this kernel), so they stay in
// the re-layout kernel (multi_pack_kernel, 7 M parameters), which the optimizer launches right after this one.
#pragma once
#include "common.cuh"

namespace b200 {

struct AdamTensor { float* p; const float* g; float* m; float* v; long n; bf16* s0; };      // one parameter (s0: bf16 mirror or null)
struct AdamChunk { int tensor; int pad; long start; };                            // kAdamChunk elements of it
static constexpr long kAdamChunk = 32768;

struct AdamHyper { float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt; const int* dev_step; };
// Capturable form (a training step replayed as a CUDA graph bakes kernel arguments in): the 1-based update count lives in device
// memory, is advanced by this one-thread kernel, and the bias corrections are formed from it inside adamw_kernel.
static __global__ void adam_step_inc_kernel(int* step) { *step += 1; }

static __global__ void __launch_bounds__(256) adamw_kernel(const AdamTensor* __restrict__ tensors, const AdamChunk* __restrict__ chunks, const AdamHyper h) {
  const AdamChunk ck = chunks[blockIdx.x];
  const AdamTensor t = tensors[ck.tensor];
  const long end = min(t.n, ck.start + kAdamChunk);
  float bc1 = h.bc1, bc2_sqrt = h.bc2_sqrt;
  if (h.dev_step) {
    const double n = (double)*h.dev_step;
    bc1 = (float)(1.0 - pow((double)h.beta1, n)); bc2_sqrt = (float)sqrt(1.0 - pow((double)h.beta2, n));
  }
  const float decay = 1.f - h.lr * h.weight_decay, step = h.lr / bc1, ob1 = 1.f - h.beta1, ob2 = 1.f - h.beta2, inv_bc2 = 1.f / bc2_sqrt;
  const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(t.s0) & 7) == 0;
  if (vec) {
    const long e4 = ck.start + ((end - ck.start) & ~3L);
#pragma unroll 2
    for (long i = ck.start + 4 * threadIdx.x; i < e4; i += 4 * 256) {
      float4 g = *reinterpret_cast<const float4*>(t.g + i), p = *reinterpret_cast<float4*>(t.p + i);
      float4 m = *reinterpret_cast<float4*>(t.m + i), v = *reinterpret_cast<float4*>(t.v + i);
      float gg[4] = {g.x, g.y, g.z, g.w}, pp[4] = {p.x, p.y, p.z, p.w}, mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pp[j] *= decay;
        mm[j] = h.beta1 * mm[j] + ob1 * gg[j];
        vv[j] = h.beta2 * vv[j] + ob2 * gg[j] * gg[j];
        pp[j] -= step * mm[j] / (sqrtf(vv[j]) * inv_bc2 + h.eps);
      }
      *reinterpret_cast<float4*>(t.p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
      *reinterpret_cast<float4*>(t.m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(t.v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      if (t.s0) {
        __nv_bfloat162 a = __floats2bfloat162_rn(pp[0], pp[1]), b = __floats2bfloat162_rn(pp[2], pp[3]);
        uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(t.s0 + i) = o;
      }
    }
    for (long i = e4 + threadIdx.x; i < end; i += 256) {
      float g = t.g[i], p = t.p[i] * decay, m = h.beta1 * t.m[i] + ob1 * g, v = h.beta2 * t.v[i] + ob2 * g * g;
      p -= step * m / (sqrtf(v) * inv_bc2 + h.eps);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
      if (t.s0) t.s0[i] = __float2bfloat16_rn(p);
    }
  } else {
    for (long i = ck.start + threadIdx.x; i < end; i += 256) {
      float g = t.g[i], p = t.p[i] * decay, m = h.beta1 * t.m[i] + ob1 * g, v = h.beta2 * t.v[i] + ob2 * g * g;
      p -= step * m / (sqrtf(v) * inv_bc2 + h.eps);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
      if (t.s0) t.s0[i] = __float2bfloat16_rn(p);
    }
  }
}

}  // namespace b200
