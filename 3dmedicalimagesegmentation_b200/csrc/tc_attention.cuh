// Fused short-sequence self-attention forward on tcgen05 / TMEM (SABlock.forward; L = 216 tokens at 96^3):
//
//   O[b, l, h*64 + d] = sum_j softmax_j(scale * Q[b,h,l,:] . K[b,h,j,:]) * V[b,h,j,d]          head_dim = 64, L <= 256
//
// One CTA per (batch, head, 128-row block of queries).  The whole key axis fits one TMEM tile (N = L rounded up to 16 <= 256
// fp32 columns), so there is no online-softmax rescaling and the scores never leave the SM:
//   warp 0      TMA: Q block [128 x 64], K [LK x 64] (both K-major, 128B swizzle), V as LK/64 boxes [64 keys x 64] (MN-major B)
//   warp 1      MMA: S = Q K^T  (4 x tcgen05.mma 128 x LK x 16)  ->  TMEM columns [0, LK)
//               ... softmax warps publish P (bf16) in shared memory ...
//               O = P V   (LK/16 x tcgen05.mma 128 x 64 x 16)   ->  TMEM columns [256, 320)
//   warps 2..9  two threads per query row (TMEM lane; warps 2..5 take the first half of the key columns, 6..9 the second): the
//               scores are read from TMEM once into registers, (max, sum) are exchanged through shared memory, P is written as
//               bf16 into the 128B-swizzled K-major A-operand layout the second MMA reads (and to global memory for the
//               backward pass); finally O -> bf16 -> att[b, l, h*64 ..].
// Replaces GEMM + softmax kernel + GEMM (3 launches, fp32 scores through L2) of exec.cuh::attention_fwd.
#pragma once
#include "tc_gemm.cuh"

namespace b200 {
namespace tc {

struct AttnParams {
  int L, LK, Lp, H, nh, tiles_m;     // LK = L rounded up to 16; Lp = row pitch of the stored probabilities
  float scale;
  bf16* P;                           // [B][nh][L][Lp]: forward = probabilities kept for the backward pass (nullptr in inference);
                                     // backward = dS (out)
  bf16* att;                         // forward: att [B*L][H]; backward: the Q third of dqkv [B*L][3H]
  int out_pitch;                     // row pitch of `att` in elements
  long long* trace;
  long long* dbg;                    // optional: CTA 0 phase stamps (clock64)
};

__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// issue one x16 TMEM load without waiting (the caller runs tcgen05.wait::ld before the first use of v)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr) : "memory");
}

static constexpr int ATT_Q_BYTES = 128 * 128, ATT_KV_BYTES = 256 * 128, ATT_P_BYTES = 4 * 128 * 128;
static constexpr int ATT_SMEM = ATT_Q_BYTES + 2 * ATT_KV_BYTES + ATT_P_BYTES + 1024 + 64 + 2 * 128 * 8;

// BWD = 0: forward as described above.
// BWD = 1: the query-row half of the backward pass with the same data flow (attention_bwd in exec.cuh):
//   "Q" = dO block, "K" = V (K-major)  ->  dP = dO V^T in TMEM
//   row op: dS = P * (dP - sum_j dP_j P_j) * scale, with the P block TMA-loaded into the very smem tile (swizzled K-major A
//           operand layout) that dS then overwrites in place
//   "V" = K (MN-major)  ->  dQ = dS K ; dS is also copied out (whole rows) for the dK = dS^T Q product.
template <int BWD>
static __global__ void __launch_bounds__(320, 1)
attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
            const __grid_constant__ CUtensorMap map_p, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_Q_BYTES;
  uint8_t* sV = sK + ATT_KV_BYTES;
  uint8_t* sP = sV + ATT_KV_BYTES;
  uint64_t* bar_qk = (uint64_t*)(sP + ATT_P_BYTES);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_p = bar_qk + 3;
  uint64_t* bar_o = bar_qk + 4;
  uint64_t* bar_pld = bar_qk + 5;               // BWD: the P block has landed in sP
  uint32_t* tmem_slot = (uint32_t*)(bar_qk + 6);
  float2* xch = (float2*)(bar_qk + 8);           // [2 halves][128 rows] (local max, local sum)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item: (batch b, head h, query block tm)
  int t = blockIdx.x;
  const int tm = t % p.tiles_m; t /= p.tiles_m;
  const int h = t % p.nh, b = t / p.nh;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_q); prefetch_tmap(&map_k); prefetch_tmap(&map_v);
    if (BWD) prefetch_tmap(&map_p);
    mbar_init(bar_pld, 1);
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 8); mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  trace_start(p.trace);
  const bool dbg = p.dbg && blockIdx.x == 0;
  if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();

  const int kboxes = (p.LK + 63) >> 6;
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_qk, (uint32_t)(128 * 128 + p.LK * 128));
      tma_load_4d(smem_u32(sQ), &map_q, bar_qk, 0, tm * 128, h, b);
      tma_load_4d(smem_u32(sK), &map_k, bar_qk, 0, 0, h, b);
      mbar_expect_tx(bar_v, (uint32_t)(kboxes * 64 * 128));
      for (int c = 0; c < kboxes; ++c) tma_load_4d(smem_u32(sV) + c * (64 * 128), &map_v, bar_v, 0, c * 64, h, b);
      if (BWD) {
        mbar_expect_tx(bar_pld, (uint32_t)(kboxes * 128 * 128));
        for (int c = 0; c < kboxes; ++c) tma_load_4d(smem_u32(sP) + c * 16384, &map_p, bar_pld, c * 64, tm * 128, h, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t hi = desc_hi(1024, 2);
    // S = Q K^T : A = Q (K-major), B = K (K-major), N = LK
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.LK >> 3) << 17) | ((128u >> 4) << 24);
    // O = P V   : A = P (K-major, written by the softmax warps), B = V (MN-major), N = 64
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    mbar_wait(bar_qk, 0);
    tc_fence_after();
    if (dbg && lane == 0) p.dbg[1] = clock64();
    if (elect_one()) {
      const uint32_t q_lo = desc_lo(smem_u32(sQ), 16), k_lo = desc_lo(smem_u32(sK), 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) umma_f16(tmem_base, desc64(q_lo + j * 2u, hi), desc64(k_lo + j * 2u, hi), idesc_s, j ? 1u : 0u);
      umma_commit(bar_s);
    }
    __syncwarp();
    mbar_wait(bar_v, 0);
    if (dbg && lane == 0) p.dbg[7] = clock64();
    mbar_wait(bar_p, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t p_lo = desc_lo(smem_u32(sP), 16), v_lo = desc_lo(smem_u32(sV), 64 * BK * 2);
      const int ksteps = p.LK >> 4;
      for (int j = 0; j < ksteps; ++j)
        umma_f16(tmem_base + 256u, desc64(p_lo + (uint32_t)(j >> 2) * (16384u >> 4) + (uint32_t)(j & 3) * 2u, hi),
                 desc64(v_lo + (uint32_t)j * (2048u >> 4), hi), idesc_o, j ? 1u : 0u);
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    // ---- softmax + epilogue: warps 2..5 own the first half of the key columns, warps 6..9 the second half; thread = query row.
    // Each thread reads its <= 128 scores from TMEM ONCE (all loads behind one wait), keeps them in registers, and the two
    // threads of a row exchange (max, sum) through shared memory.  Branch-free: scores go to the log2 domain once
    // (s2 = scale * log2 e), every exponential is one MUFU.EX2 (ex2.approx.ftz), masked columns carry -inf (ex2 = 0), and the
    // normalised probability is e * f with ONE factor f = 2^(local max - row max) / row sum per thread (no second exp pass).
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;                // TMEM lane = query row inside the block
    const int m = tm * 128 + row;                 // query index inside the (batch, head)
    const bool valid = m < p.L;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const int nch = p.LK >> 4, nch_a = (nch + 1) >> 1;
    const int ch0 = half ? nch_a : 0, myn = half ? nch - nch_a : nch_a;      // my 16-column chunks: [ch0, ch0 + myn), myn <= 8
    const int col0 = ch0 * 16;
    const float s2 = p.scale * 1.4426950408889634f;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    if (dbg && warp == 2 && lane == 0) p.dbg[2] = clock64();
    float v[128];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c < myn) tmem_ld16_issue(trow + (uint32_t)(col0 + 16 * c), v + 16 * c);
      else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[16 * c + j] = -INFINITY;
      }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint8_t* srow = sP + row * 128;
    const int sw = row & 7;
    if constexpr (BWD) {
      // v = dP (fp32, raw).  P (bf16) is read from the TMA-loaded smem block, chunk by chunk, twice (64 packed registers would not
      // fit next to v[128] at 320 threads); padding columns of P are zero-filled by the TMA, so dS is zero there.
      mbar_wait(bar_pld, 0);
      float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (u < 2 * myn) {
          const int c = col0 + 8 * u;
          const uint4 pv = *reinterpret_cast<const uint4*>(srow + (c >> 6) * 16384 + ((((c & 63) >> 3) ^ sw) << 4));
          const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pw[j]));
            d4[j] = fmaf(v[8 * u + 2 * j], pf.x, d4[j]);
            d4[j] = fmaf(v[8 * u + 2 * j + 1], pf.y, d4[j]);
          }
        }
      xch[half * 128 + row] = make_float2((d4[0] + d4[1]) + (d4[2] + d4[3]), 0.f);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float dot = xch[row].x + xch[128 + row].x;
      if (dbg && warp == 2 && lane == 0) p.dbg[3] = clock64();
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (u < 2 * myn) {
          const int c = col0 + 8 * u;
          uint4* slot = reinterpret_cast<uint4*>(srow + (c >> 6) * 16384 + ((((c & 63) >> 3) ^ sw) << 4));
          const uint4 pv = *slot;
          const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 pf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pw[j]));
            __nv_bfloat162 h2 = __floats2bfloat162_rn(pf.x * (v[8 * u + 2 * j] - dot) * p.scale, pf.y * (v[8 * u + 2 * j + 1] - dot) * p.scale);
            pk[j] = *reinterpret_cast<uint32_t*>(&h2);
          }
          *slot = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    } else {
      // columns >= L exist only in the chunk that straddles L (L is not a multiple of 16): mask that one chunk (warp-uniform branch);
      // chunks past my range already hold -inf.  max commutes with the positive scale, so the scores stay raw until the one FFMA
      // that feeds the exponential: FMNMX + FFMA + MUFU + FADD per element.
      {
        const int lim = p.L - col0;                // my columns [lim, ...) are padding
  #pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < myn && 16 * c + 16 > lim) {
  #pragma unroll
            for (int j = 0; j < 16; ++j) if (16 * c + j >= lim) v[16 * c + j] = -INFINITY;
          }
      }
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  #pragma unroll
      for (int j = 0; j < 128; ++j) m4[j & 3] = fmaxf(m4[j & 3], v[j]);
      const float mraw = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const float mxl = mraw * s2;                            // local max in the log2 domain (-inf for a half without valid columns)
      const float mref = mraw == -INFINITY ? 0.f : mxl;       // ... whose e are then all 0, without NaN
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
  #pragma unroll
      for (int j = 0; j < 128; ++j) { v[j] = ex2_ftz(fmaf(v[j], s2, -mref)); s4[j & 3] += v[j]; }
      const float suml = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      xch[half * 128 + row] = make_float2(mxl, suml);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2 o = xch[(half ^ 1) * 128 + row];
      const float mxg = fmaxf(mxl, o.x);                      // finite: the first half always holds column 0
      const float sumg = suml * ex2_ftz(mxl - mxg) + o.y * ex2_ftz(o.x - mxg);
      const float f = ex2_ftz(mref - mxg) / sumg;
      if (dbg && warp == 2 && lane == 0) p.dbg[3] = clock64();
  #pragma unroll
      for (int u = 0; u < 16; ++u) {                          // 8 columns = one 16-byte chunk per step
        if (u < 2 * myn) {
          uint32_t pk[4];
  #pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * u + 2 * j] * f, v[8 * u + 2 * j + 1] * f);
            pk[j] = *reinterpret_cast<uint32_t*>(&h2);
          }
          const int c = col0 + 8 * u;
          // K-major 128B-swizzled A operand: k-block of 64 keys = [128 rows][128 B]; 16-byte chunk ch of row r sits at ch ^ (r & 7)
          *reinterpret_cast<uint4*>(srow + (c >> 6) * 16384 + ((((c & 63) >> 3) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    // generic-proxy smem writes -> visible to the tensor core (async proxy), then hand over to the MMA warp
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    if (dbg && warp == 2 && lane == 0) p.dbg[4] = clock64();
    if (p.P) {
      // probabilities for the backward pass: copied out of the swizzled smem tile as whole rows (Lp * 2 contiguous bytes per warp
      // store; per-thread row stores were 32 sectors per instruction), while the P V MMAs run
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int wi = warp - 2;
      bf16* pbase = p.P + (((long)b * p.nh + h) * p.L) * p.Lp;
      const int nchunk = p.Lp >> 3;                          // 16-byte chunks per stored row (Lp % 8 == 0)
#pragma unroll 4
      for (int r = 0; r < 16; ++r) {
        const int rr = wi * 16 + r, mr = tm * 128 + rr;
        if (mr < p.L && lane < nchunk)
          *reinterpret_cast<uint4*>(pbase + (long)mr * p.Lp + lane * 8) =
              *reinterpret_cast<const uint4*>(sP + (lane >> 3) * 16384 + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
      }
    }
    mbar_wait(bar_o, 0);
    tc_fence_after();
    if (dbg && warp == 2 && lane == 0) p.dbg[5] = clock64();
    bf16* orow = p.att + ((long)b * p.L + m) * p.out_pitch + h * 64 + half * 32;      // each half stores 32 of the 64 output columns
    float ov[32];
    tmem_ld16x2(trow + 256u + (uint32_t)(half * 32), ov, true);
    if (valid) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { __nv_bfloat162 h2 = __floats2bfloat162_rn(ov[8 * u + 2 * j], ov[8 * u + 2 * j + 1]); pk[j] = *reinterpret_cast<uint32_t*>(&h2); }
        *reinterpret_cast<uint4*>(orow + 8 * u) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  trace_end(p.trace);
  if (dbg && threadIdx.x == 0) p.dbg[6] = clock64();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Key-row half of the attention backward in one launch (was two batched GEMM launches):
//   dV[keys, d] = P^T dO      dK[keys, d] = dS^T Q        per (sample, head, 128-key block), reduction over the <= 256 queries
// Both products read their A operand transposed out of a [queries][keys] row-major matrix (MN-major A: 64-key x 64-query TMA boxes)
// and their B operand as [queries][64] (MN-major B); two fp32 accumulators of 64 columns sit side by side in TMEM.
struct AttnKvParams {
  int L, LK, H, nh, tiles_k;
  bf16* dqkv;                        // [B*L][3H]: dK goes to the K third, dV to the V third
  long long* trace;
};
static constexpr int ATT_KV_SMEM = 2 * 4 * 16384 + 2 * 4 * 8192 + 1024 + 256;

static __global__ void __launch_bounds__(192, 1)
attn_bwd_kv_kernel(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_do, const __grid_constant__ CUtensorMap map_ds,
                   const __grid_constant__ CUtensorMap map_q, const AttnKvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA1 = smem;                 // P^T  : 4 query blocks x (2 boxes of [64 queries][64 keys])
  uint8_t* sA2 = sA1 + 4 * 16384;      // dS^T
  uint8_t* sB1 = sA2 + 4 * 16384;      // dO   : 4 query blocks x [64 queries][64]
  uint8_t* sB2 = sB1 + 4 * 8192;       // Q
  uint64_t* bar1 = (uint64_t*)(sB2 + 4 * 8192);
  uint64_t* bar2 = bar1 + 1;
  uint64_t* bar_done = bar1 + 2;
  uint32_t* tmem_slot = (uint32_t*)(bar1 + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int tk = t % p.tiles_k; t /= p.tiles_k;
  const int h = t % p.nh, b = t / p.nh;
  const int m0 = tk * 128;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_p); prefetch_tmap(&map_do); prefetch_tmap(&map_ds); prefetch_tmap(&map_q);
    mbar_init(bar1, 1); mbar_init(bar2, 1); mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  trace_start(p.trace);
  const int kblocks = (p.LK + 63) >> 6;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar1, (uint32_t)(kblocks * (16384 + 8192)));
      for (int kb = 0; kb < kblocks; ++kb) {
        tma_load_4d(smem_u32(sA1) + kb * 16384, &map_p, bar1, m0, kb * 64, h, b);
        tma_load_4d(smem_u32(sA1) + kb * 16384 + 8192, &map_p, bar1, m0 + 64, kb * 64, h, b);
        tma_load_4d(smem_u32(sB1) + kb * 8192, &map_do, bar1, 0, kb * 64, h, b);
      }
      mbar_expect_tx(bar2, (uint32_t)(kblocks * (16384 + 8192)));
      for (int kb = 0; kb < kblocks; ++kb) {
        tma_load_4d(smem_u32(sA2) + kb * 16384, &map_ds, bar2, m0, kb * 64, h, b);
        tma_load_4d(smem_u32(sA2) + kb * 16384 + 8192, &map_ds, bar2, m0 + 64, kb * 64, h, b);
        tma_load_4d(smem_u32(sB2) + kb * 8192, &map_q, bar2, 0, kb * 64, h, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // both operands MN-major (bits 15, 16), N = 64, M = 128; A: 64-wide M atoms 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO),
    // +2048 B per 16-wide K step; B likewise with a single 64-wide N atom
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi = desc_hi(1024, 2);
    const int ksteps = p.LK >> 4;
    for (int prod = 0; prod < 2; ++prod) {
      mbar_wait(prod ? bar2 : bar1, 0);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = desc_lo(smem_u32(prod ? sA2 : sA1), 8192), b_lo = desc_lo(smem_u32(prod ? sB2 : sB1), 8192);
        for (int j = 0; j < ksteps; ++j)
          umma_f16(tmem_base + (uint32_t)(prod * 64),
                   desc64(a_lo + (uint32_t)(j >> 2) * (16384u >> 4) + (uint32_t)(j & 3) * (2048u >> 4), hi),
                   desc64(b_lo + (uint32_t)(j >> 2) * (8192u >> 4) + (uint32_t)(j & 3) * (2048u >> 4), hi), idesc, j ? 1u : 0u);
        if (prod) umma_commit(bar_done);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int key = m0 + q * 32 + lane;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
    for (int prod = 0; prod < 2; ++prod) {
      // prod 0 = dV -> V third (column offset 2H), prod 1 = dK -> K third (offset H)
      bf16* orow = p.dqkv + ((long)b * p.L + key) * (3 * p.H) + (prod ? p.H : 2 * p.H) + h * 64;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float v[32];
        tmem_ld16x2(trow + (uint32_t)(prod * 64 + c0), v, true);
        if (key < p.L) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * u + 2 * j], v[8 * u + 2 * j + 1]); pk[j] = *reinterpret_cast<uint32_t*>(&h2); }
            *reinterpret_cast<uint4*>(orow + c0 + 8 * u) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  trace_end(p.trace);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

static inline bool attention_fused_supported(int L, int Lp, int H, int nh) {
  static const bool off = getenv("B200_NO_FUSED_ATTENTION") != nullptr;
  return !off && nh > 0 && H / nh == 64 && H % 8 == 0 && L >= 16 && ((L + 15) & ~15) <= 256 && Lp % 8 == 0 && Lp >= L;
}

// qkv: [B*L][3H] bf16, columns [Q | K | V], head h at h*64 inside each third.  P (optional): [B][nh][L][Lp].  att: [B*L][H].
static int attention_fused_fwd(const bf16* qkv, bf16* P, bf16* att, int B, int nh, int L, int Lp, int H, float scale, cudaStream_t st) {
  const int dh = 64;
  const long sQb = (long)L * 3 * H;
  AttnParams p;
  p.L = L; p.LK = (L + 15) & ~15; p.Lp = Lp; p.H = H; p.nh = nh; p.tiles_m = cdiv(L, 128); p.scale = scale; p.P = P; p.att = att; p.out_pitch = H;
  CUtensorMap mq, mk, mv;
  B200_TRY(make_map(&mq, operand(qkv, 3 * H, 1, sQb, dh), L, dh, 128, B, nh));
  B200_TRY(make_map(&mk, operand(qkv + H, 3 * H, 1, sQb, dh), L, dh, p.LK, B, nh));
  B200_TRY(make_map(&mv, operand(qkv + 2 * H, 1, 3 * H, sQb, dh), dh, L, 64, B, nh));
  static bool attr_done = false;
  if (!attr_done) { B200_CUDA(cudaFuncSetAttribute(attn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM)); attr_done = true; }
  p.trace = trace_slot(); if (p.trace) trace_tag("attn_fwd L%d b%d", L, B * nh);
  p.dbg = g_dbg;
  B200_CUDA(launch_pdl(attn_kernel<0>, dim3(B * nh * p.tiles_m), dim3(320), (size_t)ATT_SMEM, st, mq, mk, mv, mq, p));
  B200_LAUNCH_CHECK();
  return 0;
}

// Backward, query-row half: dS = P * (dO V^T - rowsum(dO V^T * P)) * scale -> dS [B][nh][L][Lp] and dQ = dS K -> dqkv[:, h*64 ..] (Q third
// of the [B*L][3H] gradient).  datt: [B*L][H] gradient of the attention output; P: probabilities of the forward.
static int attention_fused_bwd_dq(const bf16* qkv, const bf16* P, const bf16* datt, bf16* dS, bf16* dqkv, int B, int nh, int L, int Lp, int H,
                                  float scale, cudaStream_t st) {
  const int dh = 64;
  const long sQb = (long)L * 3 * H, sOb = (long)L * H, sPb = (long)nh * L * Lp, sPh = (long)L * Lp;
  AttnParams p;
  p.L = L; p.LK = (L + 15) & ~15; p.Lp = Lp; p.H = H; p.nh = nh; p.tiles_m = cdiv(L, 128); p.scale = scale; p.P = dS; p.att = dqkv; p.out_pitch = 3 * H;
  CUtensorMap mq, mk, mv, mp;
  B200_TRY(make_map(&mq, operand(datt, H, 1, sOb, dh), L, dh, 128, B, nh));                 // "Q" = dO block
  B200_TRY(make_map(&mk, operand(qkv + 2 * H, 3 * H, 1, sQb, dh), L, dh, p.LK, B, nh));     // "K" = V, K-major
  B200_TRY(make_map(&mv, operand(qkv + H, 1, 3 * H, sQb, dh), dh, L, 64, B, nh));           // "V" = K, MN-major
  B200_TRY(make_map(&mp, operand(P, Lp, 1, sPb, sPh), L, L, 128, B, nh));                   // P block: 64-column boxes of 128 rows
  static bool attr_done = false;
  if (!attr_done) { B200_CUDA(cudaFuncSetAttribute(attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM)); attr_done = true; }
  p.trace = trace_slot(); if (p.trace) trace_tag("attn_bwd_dq L%d b%d", L, B * nh);
  p.dbg = g_dbg;
  B200_CUDA(launch_pdl(attn_kernel<1>, dim3(B * nh * p.tiles_m), dim3(320), (size_t)ATT_SMEM, st, mq, mk, mv, mp, p));
  B200_LAUNCH_CHECK();
  return 0;
}

// Backward, key-row half: dV = P^T dO -> V third of dqkv, dK = dS^T Q -> K third.
static int attention_fused_bwd_kv(const bf16* qkv, const bf16* P, const bf16* dS, const bf16* datt, bf16* dqkv, int B, int nh, int L, int Lp, int H,
                                  cudaStream_t st) {
  const int dh = 64;
  const long sQb = (long)L * 3 * H, sOb = (long)L * H, sPb = (long)nh * L * Lp, sPh = (long)L * Lp;
  AttnKvParams p;
  p.L = L; p.LK = (L + 15) & ~15; p.H = H; p.nh = nh; p.tiles_k = cdiv(L, 128); p.dqkv = dqkv;
  CUtensorMap mp, mdo, mds, mq;
  B200_TRY(make_map(&mp, operand(P, 1, Lp, sPb, sPh), L, L, 128, B, nh));             // A = P^T   (MN-major: keys contiguous)
  B200_TRY(make_map(&mdo, operand(datt, 1, H, sOb, dh), dh, L, 64, B, nh));           // B = dO    (MN-major: d contiguous)
  B200_TRY(make_map(&mds, operand(dS, 1, Lp, sPb, sPh), L, L, 128, B, nh));           // A = dS^T
  B200_TRY(make_map(&mq, operand(qkv, 1, 3 * H, sQb, dh), dh, L, 64, B, nh));         // B = Q
  static bool attr_done = false;
  if (!attr_done) { B200_CUDA(cudaFuncSetAttribute(attn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_KV_SMEM)); attr_done = true; }
  p.trace = trace_slot(); if (p.trace) trace_tag("attn_bwd_kv L%d b%d", L, B * nh);
  B200_CUDA(launch_pdl(attn_bwd_kv_kernel, dim3(B * nh * p.tiles_k), dim3(192), (size_t)ATT_KV_SMEM, st, mp, mdo, mds, mq, p));
  B200_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc
}  // namespace b200
