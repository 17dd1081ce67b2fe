// tcgen05 3x3x3 convolution for the full-resolution layers (16 / 32 output channels): the three kw taps of a (kd, kh) pair are
// STACKED ALONG N, so one MMA is 128 x 3*Co x 16 instead of three MMAs of 128 x Co x 16.
//
// Why: a tcgen05.mma with a 128-row A tile costs ~60 cycles of A-operand smem read whether N is 16 or 48 (measured);
// tc_conv_halo.cuh issues 27 of them per 128 output voxels (16 -> 16 @96^3: 78 us against an 18 us roofline).  Here
//
//     P[(h, w'), (kw, co)] = sum_{kd, kh, ci} halo[d + kd][h + kh][w'][ci] * W[kd][kh][kw][co][ci]          w' = 0 .. 7 (halo columns)
//     out[h][w][co]        = P[(h, w), (0, co)] + P[(h, w + 1), (1, co)] + P[(h, w + 2), (2, co)]            w  = 0 .. 5
//
// The output tile is 16 (h) x 6 (w) voxels of one d-plane; its halo plane is 18 x 8 voxels, so a halo LINE is exactly one 8-row
// UMMA group: the A operand of (kd, kh) is the plain contiguous K-major view that starts kh lines into plane d + kd (SBO = 8 rows),
// and the three kw weight matrices are 3*Co consecutive rows of the packed [tap][co][ci] weights -- no re-packing.  9 MMAs per
// k-step and output plane instead of 27; 96 of the 128 accumulator rows are outputs (the other 32 are the two halo columns).  The
// shift-add runs in the epilogue: a TMEM lane is a (h, w') row, w' = lane % 8, so P(.., w'+1) and P(.., w'+2) are shfl_down 1 and 2
// inside an 8-lane group.
//
// Same pipeline as tc_conv_halo.cuh (one 5-D TMA box of op + 2 halo planes per tile, resident weights, double-buffered TMEM, `op`
// output planes per tile) and the same fused 1x1x1 convolution of the residual block:
//   mode2 == 1 (forward):  second output = centre-tap view (one line + one column in) x W3, N = Co, own accumulator columns
//   mode2 == 2 (dgrad):    second input tile (16 x 8 voxels, columns 6, 7 unused) x W3^T into the kw = 0 columns
// Used when Ci in {16, 32, 64} (one channel chunk) and Co = 16 (the full-resolution layers); everything else stays on tc_conv_halo.cuh.
// Measured on B200, configs[1] (in-situ trace, us per launch): 16 -> 16 @96^3 77.6 -> 61.3, 32 -> 16 @96^3 with the fused 1x1x1 154 -> 88.
// ncu (profiles/r02_summary.md): the tensor sub-pipe is busy 87 % of the kernel, 96 cycles per 128 x 48 x 16 MMA.
#pragma once
#include "tc_conv.cuh"

namespace b200 {
namespace tc {

static constexpr int S48_TH = 16, S48_TW = 6, S48_HW = 8, S48_HH = 18;

struct Halo48Params {
  int N, D, H, W, Ci, Co;
  int row_bytes, plane_bytes, halo_bytes, x2_off, stage_bytes, stages, w_bytes;
  int tiles_w, tiles_h, tiles_per_n, total_tiles, op;
  int acc_cols; uint32_t tmem_cols;
  bf16* out; int pitch, coff, accumulate; double* stats; int out_half;
  int mode2; bf16* out2; int pitch2, coff2; double* stats2;
  // Backward-norm sums in the dgrad epilogue (stats != null, gact != null): the output IS the gradient wrt a = lrelu(norm(c)); with the
  // saved activation row the epilogue accumulates Sg = sum g and Sgn = sum g*n (g = out * lrelu'(a), n recovered from a) instead of
  // sum x / sum x^2 -- the first pass of the InstanceNorm backward (in_bwd_reduce_kernel<false>) without reading the tensor again.
  const bf16* gact; int gact_pitch, gact_coff; int stats_stride;
  long long* trace;
};

template <int KSTEPS, int CO_T>
__global__ void __launch_bounds__(192, 1)
conv_halo48_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x2,
                   const __grid_constant__ CUtensorMap map_w2, const Halo48Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem + p.w_bytes;
  uint64_t* full = (uint64_t*)(ring + (size_t)p.stages * p.stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(wfull + 1);
  float* red = (float*)(tmem_slot + 4);
  constexpr uint32_t ROW_BYTES = 32 * KSTEPS, ROW_UNITS = 2 * KSTEPS;      // smem row of one voxel: Ci bf16

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  trace_start(p.trace);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // resident weights: 9 boxes of 3*Co rows (the kw triple of one (kd, kh)), then the fused 1x1x1 matrix
    const uint32_t wb = (uint32_t)(3 * p.Co) * ROW_BYTES;
    mbar_expect_tx(wfull, 9u * wb + (p.mode2 ? (uint32_t)p.Co * ROW_BYTES : 0u));
    for (int g = 0; g < 9; ++g) tma_load_2d(smem_u32(smem) + (uint32_t)g * wb, &map_w, wfull, 0, g * 3 * p.Co);
    if (p.mode2) tma_load_2d(smem_u32(smem) + 9u * wb, &map_w2, wfull, 0, 0);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // (the resident weights fetched above were packed many launches earlier)

  if (warp == 0) {
    // ---- TMA producer: one halo box (op + 2 planes of 18 x 8 voxels) per tile (+ the second input tile of the fused dgrad)
    int stage = 0; uint32_t phase = 0;
    const uint32_t ring_u = smem_u32(ring);
    const uint32_t tx = (uint32_t)((p.op + 2) * p.plane_bytes) + (p.mode2 == 2 ? (uint32_t)(p.op * 128) * ROW_BYTES : 0u);
    int t = blockIdx.x;
    int n = t / p.tiles_per_n, r = t - n * p.tiles_per_n;
    for (; t < p.total_tiles; t += gridDim.x) {
      const int tw = r % p.tiles_w, q = r / p.tiles_w, th = q % p.tiles_h, d = (q / p.tiles_h) * p.op;
      mbar_wait(empty + stage, phase ^ 1);
      if (elect_one()) {
        const uint32_t base = ring_u + (uint32_t)stage * p.stage_bytes;
        mbar_expect_tx(full + stage, tx);
        tma_load_5d(base, &map_x, full + stage, 0, tw * S48_TW - 1, th * S48_TH - 1, d - 1, n);
        if (p.mode2 == 2) tma_load_5d(base + p.x2_off, &map_x2, full + stage, 0, tw * S48_TW, th * S48_TH, d, n);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
      r += gridDim.x;
      while (r >= p.tiles_per_n) { r -= p.tiles_per_n; ++n; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: 9 MMAs of 128 x 3*Co x 16 per k-step and output plane
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((3 * p.Co) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Co >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t layout = KSTEPS == 4 ? 2u : (KSTEPS == 2 ? 4u : 6u);
    const uint32_t hi = desc_hi(8 * ROW_BYTES, layout);                       // contiguous 8-row groups, A and B alike
    const uint32_t a_lo0 = desc_lo(smem_u32(ring), 16), b_lo0 = desc_lo(smem_u32(smem), 16);
    const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4, plane_units = (uint32_t)p.plane_bytes >> 4;
    const uint32_t b_grp = (uint32_t)(3 * p.Co) * ROW_UNITS, b_k1 = b_lo0 + 9u * b_grp;
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    mbar_wait(wfull, 0);
    tc_fence_after();
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait(tempty + acc, acc_phase ^ 1);
      tc_fence_after();
      mbar_wait(full + stage, phase);
      tc_fence_after();
      const uint32_t a_st = a_lo0 + (uint32_t)stage * stage_units;
      if (elect_one()) {
        for (int o = 0; o < p.op; ++o) {
          const uint32_t tmem_d = tmem_base + (uint32_t)((acc * p.op + o) * p.acc_cols);
#pragma unroll
          for (int kd = 0; kd < 3; ++kd)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                const uint32_t a_lo = a_st + (uint32_t)(o + kd) * plane_units + (uint32_t)(kh * S48_HW) * ROW_UNITS + 2 * k;
                const uint32_t b_lo = b_lo0 + (uint32_t)(kd * 3 + kh) * b_grp + 2 * k;
                umma_f16(tmem_d, desc64(a_lo, hi), desc64(b_lo, hi), idesc, (kd | kh | k) ? 1u : 0u);
              }
          if (p.mode2 == 1) {        // centre tap: one halo line + one column in; rows (h, w') -> out2(h, w') for w' < 6
            const uint32_t a_c = a_st + (uint32_t)(o + 1) * plane_units + (uint32_t)(S48_HW + 1) * ROW_UNITS;
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
              umma_f16(tmem_d + (uint32_t)(3 * p.Co), desc64(a_c + 2 * k, hi), desc64(b_k1 + 2 * k, hi), idesc1, k ? 1u : 0u);
          } else if (p.mode2 == 2) { // second input tile (no halo) into the kw = 0 columns: rows (h, w') <-> output (h, w')
            const uint32_t a_2 = a_st + ((uint32_t)p.x2_off >> 4) + (uint32_t)o * (128u * ROW_UNITS);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k)
              umma_f16(tmem_d, desc64(a_2 + 2 * k, hi), desc64(b_k1 + 2 * k, hi), idesc1, 1u);
          }
        }
        umma_commit(empty + stage);
        umma_commit(tfull + acc);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ---- epilogue: TMEM lane = (h = 4 * quarter + lane / 8, w' = lane % 8); shift-add over kw, then as tc_conv_halo.cuh
    const int q4 = warp & 3, ew = warp - 2;
    const int wq = lane & 7, hq = q4 * 4 + (lane >> 3);
    const int npass = p.mode2 == 1 ? 2 : 1;
    int acc = 0; uint32_t acc_phase = 0;
    int t = blockIdx.x;
    int n = t / p.tiles_per_n, r = t - n * p.tiles_per_n;
    float rs1[2][CO_T], rs2[2][CO_T];
#pragma unroll
    for (int j = 0; j < CO_T; ++j) { rs1[0][j] = rs2[0][j] = rs1[1][j] = rs2[1][j] = 0.f; }
    int n_acc = n;
    auto flush = [&](int nn) {       // all 128 epilogue threads
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        double* sp = ps ? p.stats2 : p.stats;
        if (ps >= npass || !sp) continue;
#pragma unroll
        for (int j = 0; j < CO_T; ++j) {
          float s1 = warp_sum(rs1[ps][j]), s2 = warp_sum(rs2[ps][j]);
          if (lane == 0) { red[ew * 2 * CO_T + j] = s1; red[ew * 2 * CO_T + CO_T + j] = s2; }
          rs1[ps][j] = 0.f; rs2[ps][j] = 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int e = ew * 32 + lane;
        if (e < 2 * CO_T) {
          float tot = red[e] + red[2 * CO_T + e] + red[4 * CO_T + e] + red[6 * CO_T + e];
          int c = e % CO_T, which = e / CO_T;
          atomicAdd(sp + ((long)nn * CO_T + c) * p.stats_stride + which, (double)tot);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    };
    for (; t < p.total_tiles; t += gridDim.x) {
      const int tw = r % p.tiles_w, qq = r / p.tiles_w, th = qq % p.tiles_h, d0 = (qq / p.tiles_h) * p.op;
      if (n != n_acc) { flush(n_acc); n_acc = n; }
      const int w = tw * S48_TW + wq, h = th * S48_TH + hq;
      const bool valid = (wq < S48_TW) && (w < p.W) && (h < p.H);
      mbar_wait(tfull + acc, acc_phase);
      tc_fence_after();
      for (int o = 0; o < p.op; ++o) {
        const int d = d0 + o;
        if (d >= p.D) break;
        const long vox = (((long)n * p.D + d) * p.H + h) * p.W + w;
        const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((acc * p.op + o) * p.acc_cols);
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          if (ps >= npass) continue;
          bf16* dst = ps ? p.out2 + vox * p.pitch2 + p.coff2 : p.out + vox * p.pitch + p.coff;
          double* sp = ps ? p.stats2 : p.stats;
#pragma unroll
          for (int c0 = 0; c0 < CO_T; c0 += 16) {
            float v[16];
            if (ps == 0) {
              float v1[16], v2[16];
              tmem_ld16(trow + c0, v); tmem_ld16(trow + CO_T + c0, v1); tmem_ld16(trow + 2 * CO_T + c0, v2);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                v[j] = (v[j] + __shfl_down_sync(0xffffffffu, v1[j], 1)) + __shfl_down_sync(0xffffffffu, v2[j], 2);
            } else {
              tmem_ld16(trow + 3 * CO_T + c0, v);
            }
            if (valid) {
              if (sp) {
                if (p.gact && ps == 0) {
                  Vec16<bf16> a0, a1;
                  const bf16* ar = p.gact + vox * p.gact_pitch + p.gact_coff + c0;
                  a0.load(ar); a1.load(ar + 8);
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const float a = j < 8 ? a0.v[j] : a1.v[j - 8];
                    const float g = v[j] * (a > 0.f ? 1.f : 0.01f), nn = a > 0.f ? a : a * 100.f;
                    rs1[ps][c0 + j] += g; rs2[ps][c0 + j] = fmaf(g, nn, rs2[ps][c0 + j]);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) { rs1[ps][c0 + j] += v[j]; rs2[ps][c0 + j] = fmaf(v[j], v[j], rs2[ps][c0 + j]); }
                }
              }
              if (p.accumulate && ps == 0) {
                Vec16<bf16> a, b; a.load(dst + c0); b.load(dst + c0 + 8);
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] += a.v[j]; v[8 + j] += b.v[j]; }
              }
              if (p.out_half) {
                Vec16<__half> o0, o1;
#pragma unroll
                for (int j = 0; j < 8; ++j) { o0.v[j] = v[j]; o1.v[j] = v[8 + j]; }
                o0.store(reinterpret_cast<__half*>(dst) + c0); o1.store(reinterpret_cast<__half*>(dst) + c0 + 8);
              } else {
                Vec16<bf16> o0, o1;
#pragma unroll
                for (int j = 0; j < 8; ++j) { o0.v[j] = v[j]; o1.v[j] = v[8 + j]; }
                o0.store(dst + c0); o1.store(dst + c0 + 8);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      r += gridDim.x;
      while (r >= p.tiles_per_n) { r -= p.tiles_per_n; ++n; }
    }
    flush(n_acc);
  }
  tc_fence_before();
  __syncthreads();
  trace_end(p.trace);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

struct Halo48Plan { int rb, plane_bytes, halo_bytes, x2_bytes, stage_bytes, stages, w_bytes, op, acc_cols; uint32_t tmem_cols; bool ok; };
static inline Halo48Plan halo48_plan(int Ci, int Co, int mode2, int D, long tiles_hw_n) {
  Halo48Plan h; memset(&h, 0, sizeof(h));
  // Co = 32 (N = 96) was measured no faster than the per-tap kernel (16 -> 32 @96^3 dgrad: 106 vs 87 us): ncu shows the tensor pipe busy
  // ~42 + 1.1 * N cycles per 128-row MMA at these 32..64-byte operand rows (smem operand fetch), so stacking only pays while N is small.
  // B200_HALO48_CO32=1 switches it on for experiments.
  static const bool co32 = getenv("B200_HALO48_CO32") != nullptr;
  if (!((Ci == 16 || Ci == 32 || Ci == 64) && (Co == 16 || (Co == 32 && co32)))) return h;
  static const bool off = getenv("B200_NO_HALO48") != nullptr;
  if (off) return h;
  h.rb = Ci * 2; h.plane_bytes = S48_HH * S48_HW * h.rb;
  h.w_bytes = (((27 + (mode2 ? 1 : 0)) * Co * h.rb + 1023) / 1024) * 1024;
  h.acc_cols = 3 * Co + (mode2 == 1 ? Co : 0);
  const int budget = 200 * 1024 - h.w_bytes;
  const int op_max = getenv("B200_HALO_OP") ? atoi(getenv("B200_HALO_OP")) : 4;
  const long min_tiles = getenv("B200_HALO_MIN_TILES") ? atol(getenv("B200_HALO_MIN_TILES")) : 148 * 8;
  for (int op = 4; op >= 1; op >>= 1) {
    if (op > op_max) continue;
    if (op > 1 && (D <= 0 || D % op != 0 || (long)(D / op) * tiles_hw_n < min_tiles)) continue;
    if (2 * op * h.acc_cols > 512) continue;
    const int hb = (((op + 2) * h.plane_bytes + 1023) / 1024) * 1024, xb = mode2 == 2 ? ((op * 128 * h.rb + 1023) / 1024) * 1024 : 0;
    if (budget / (hb + xb) < (op > 1 ? 3 : 2)) continue;
    h.op = op; h.halo_bytes = hb; h.x2_bytes = xb; h.stage_bytes = hb + xb;
    h.stages = budget / h.stage_bytes; if (h.stages > 8) h.stages = 8;
    uint32_t cols = 2u * op * h.acc_cols, pw = 32; while (pw < cols) pw <<= 1; h.tmem_cols = pw;
    h.ok = true;
    break;
  }
  return h;
}

template <int KSTEPS, int CO_T>
static int conv_halo48_launch(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& mx2, const CUtensorMap& mw2, const Halo48Params& p, int grid,
                              size_t smem, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) { B200_CUDA(cudaFuncSetAttribute(conv_halo48_kernel<KSTEPS, CO_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_done = true; }
  B200_CUDA(launch_pdl(conv_halo48_kernel<KSTEPS, CO_T>, dim3(grid), dim3(192), smem, st, mx, mw, mx2, mw2, p));
  B200_LAUNCH_CHECK();
  return 0;
}

// same contract as tc::conv_halo (tc_conv_halo.cuh); mode2 / wp2 / out2 / x2 describe the fused 1x1x1 convolution
static int conv_halo48(const Halo48Plan& h, const bf16* x, int in_pitch, int in_coff, int Ci, int N, int D, int H, int W, const bf16* wp, int Co,
                       bf16* out, int out_pitch, int out_coff, int accumulate, double* stats, cudaStream_t st, int mode2, const bf16* wp2,
                       bf16* out2, int pitch2, int coff2, double* stats2, const bf16* x2, int x2_pitch, int x2_coff, int out_half,
                       const bf16* gact = nullptr, int gact_pitch = 0, int gact_coff = 0) {
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
  Halo48Params p; memset(&p, 0, sizeof(p));
  p.N = N; p.D = D; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co;
  p.row_bytes = h.rb; p.plane_bytes = h.plane_bytes; p.halo_bytes = h.halo_bytes; p.x2_off = h.halo_bytes; p.stage_bytes = h.stage_bytes;
  p.stages = h.stages; p.w_bytes = h.w_bytes; p.op = h.op; p.acc_cols = h.acc_cols; p.tmem_cols = h.tmem_cols;
  p.tiles_w = cdiv(W, S48_TW); p.tiles_h = cdiv(H, S48_TH);
  p.tiles_per_n = cdiv(D, h.op) * p.tiles_h * p.tiles_w;
  const long total = (long)N * p.tiles_per_n;
  B200_CHECK(total < (1L << 30), "halo conv: too many tiles");
  p.total_tiles = (int)total;
  p.out = out; p.pitch = out_pitch; p.coff = out_coff; p.accumulate = accumulate; p.stats = stats; p.out_half = out_half;
  p.mode2 = mode2; p.out2 = out2; p.pitch2 = pitch2; p.coff2 = coff2; p.stats2 = stats2;
  p.gact = gact; p.gact_pitch = gact_pitch; p.gact_coff = gact_coff; p.stats_stride = gact ? 3 : 2;
  p.trace = trace_slot(); if (p.trace) trace_tag("conv_halo48 %d->%d @%d mode2=%d", Ci, Co, D, mode2);
  const CUtensorMapSwizzle sw = h.rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (h.rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap mx, mw, mx2, mw2;
  memset(&mx2, 0, sizeof(mx2)); memset(&mw2, 0, sizeof(mw2));
  cuuint32_t es5[5] = {1, 1, 1, 1, 1}, es2[2] = {1, 1};
  {
    cuuint64_t dims[5] = {(cuuint64_t)Ci, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)W * in_pitch * 2, (cuuint64_t)H * W * in_pitch * 2, (cuuint64_t)D * H * W * in_pitch * 2};
    cuuint32_t box[5] = {(cuuint32_t)Ci, S48_HW, S48_HH, (cuuint32_t)(h.op + 2), 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(x + in_coff), dims, strides, box, es5, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "halo48 input tensor map failed (%d)", (int)r);
  }
  if (mode2 == 2) {
    cuuint64_t dims[5] = {(cuuint64_t)Ci, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)x2_pitch * 2, (cuuint64_t)W * x2_pitch * 2, (cuuint64_t)H * W * x2_pitch * 2, (cuuint64_t)D * H * W * x2_pitch * 2};
    cuuint32_t box[5] = {(cuuint32_t)Ci, S48_HW, S48_TH, (cuuint32_t)h.op, 1};
    CUresult r = enc(&mx2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(x2 + x2_coff), dims, strides, box, es5, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "halo48 second-input tensor map failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Ci, (cuuint64_t)27 * Co};
    cuuint64_t strides[1] = {(cuuint64_t)Ci * 2};
    cuuint32_t box[2] = {(cuuint32_t)Ci, (cuuint32_t)(3 * Co)};
    CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)wp, dims, strides, box, es2, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "halo48 weight tensor map failed (%d)", (int)r);
  }
  if (mode2) {
    cuuint64_t dims[2] = {(cuuint64_t)Ci, (cuuint64_t)Co};
    cuuint64_t strides[1] = {(cuuint64_t)Ci * 2};
    cuuint32_t box[2] = {(cuuint32_t)Ci, (cuuint32_t)Co};
    CUresult r = enc(&mw2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)wp2, dims, strides, box, es2, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "halo48 1x1x1 weight tensor map failed (%d)", (int)r);
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + p.w_bytes + 1024 + 256 + 8 * Co * sizeof(float) + 64;
  B200_CHECK(smem <= 227 * 1024, "halo48 smem budget exceeded (%zu)", smem);
  const int grid = (int)(p.total_tiles < num_sms() ? p.total_tiles : num_sms());
  const int ks = Ci / 16;
#define B200_H48_CASE(KS, CO) if (ks == KS && Co == CO) return conv_halo48_launch<KS, CO>(mx, mw, mx2, mw2, p, grid, smem, st)
  B200_H48_CASE(1, 16); B200_H48_CASE(1, 32); B200_H48_CASE(2, 16); B200_H48_CASE(2, 32); B200_H48_CASE(4, 16); B200_H48_CASE(4, 32);
#undef B200_H48_CASE
  B200_CHECK(false, "halo48: no kernel for Ci=%d Co=%d", Ci, Co);
}

}  // namespace tc
}  // namespace b200
