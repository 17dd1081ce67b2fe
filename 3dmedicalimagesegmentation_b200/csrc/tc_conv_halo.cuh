// tcgen05 3x3x3 convolution with on-SM halo reuse (forward and dgrad), channels-last bf16.
//
// tc_conv.cuh fetches the voxel brick once per tap: every input voxel crosses L2->SM 27 times and the kernel is bound by
// L2 request bandwidth (measured 16-35 B/clk/SM).  Here the output tile is ONE d-plane of 16(h) x 8(w) voxels and, per
// kd, one halo plane of (16+2) x (8+2) voxels [18][10][C] sits in swizzled smem.  All nine (kh,kw) taps of that kd
// are then just *descriptor views* of the same tile:
//       start = tile + (kh*10 + kw) * row_bytes ,   SBO (next 8-row group = next h) = 10 * row_bytes
// which is legal because UMMA (like TMA) applies the 32/64/128-byte swizzle XOR to absolute smem address bits (the same
// property the +32 B K-advance inside a swizzle atom relies on).  When the packed weights fit in smem next to the ring,
// ONE TMA box of depth 3 brings all three halo planes of a tile (1 TMA load and 27*kc/16 MMAs per pipeline stage).
//
// Measured on B200 (ncu, 16->16 @96^3): the tensor pipe is busy 32 cycles per 128x16x16 MMA (A-operand smem read), so the
// single issuing thread must spend well under that per MMA.  The issue loop is therefore a template on kc/16 with every
// tap offset a compile-time constant: 2 uniform adds + 1 UTCHMMA per MMA, all operands in uniform registers.
// InstanceNorm statistics of 16/32-channel outputs are accumulated in registers across a CTA's tiles and flushed with one
// warp reduction per sample instead of 2*Co shuffled reductions per tile.
#pragma once
#include "tc_conv.cuh"
#include "tc_conv_halo48.cuh"

namespace b200 {
namespace tc {

static constexpr int HTH = 16, HTW = 8, HALO_H = HTH + 2, HALO_W = HTW + 2;

struct HaloParams {
  int N, D, H, W, Ci, Co;
  int kc, row_bytes, nchunk;     // channels per chunk (<= 64), smem row size, Ci/kc
  int tiles_w, tiles_h, tiles_per_n, total_tiles;
  int planes;                    // halo planes per pipeline stage: 3 (one TMA box of depth op+2 serving `op` output planes) or 1
  int op;                        // output d-planes per tile (planes == 3 only; 1, 2 or 4): (op+2)/op halo planes fetched per output plane
  int plane_bytes;               // HALO_H*HALO_W*row_bytes
  int halo_bytes;                // planes*plane_bytes rounded up to 1024 (weight tiles of a non-resident stage follow)
  int b_bytes;                   // one (tap, chunk) weight tile, rounded to 1024
  int resident;                  // all 27*nchunk weight tiles stay in smem
  int stage_bytes, stages; uint32_t tmem_cols;
  bf16* out; int pitch, coff, accumulate; double* stats;
  // fused 1x1x1 convolution of the residual block (weight tile index 27*nchunk + chunk, resident weights only):
  //  mode2 == 1: second OUTPUT  out2 = conv1x1(x; W3)      (forward: conv1 and conv3 read the same x)
  //  mode2 == 2: second INPUT   out += conv1x1(x2; W3^T)   (dgrad: dx = dgrad3x3(dc1) + dgrad1x1(dc3))
  int out_half;   // raw conv outputs (forward, feeding an InstanceNorm) are stored as fp16 (RawOf<bf16>)
  int mode2; bf16* out2; int pitch2, coff2; double* stats2; int x2_off;
  long long* trace;
  int dbg_mode; long long* dbg;  // tuning aids: bit0 skip TMA, bit1 skip MMA, bit2 skip epilogue stores; CTA-0 clock stamps
};

// all 9 (kh,kw) taps x KSTEPS k-steps of one halo plane: a_pl/b_pl are descriptor low words (16-byte units)
template <int KSTEPS>
__device__ __forceinline__ void issue_plane(uint32_t tmem_d, uint32_t a_pl, uint32_t a_hi, uint32_t b_pl, uint32_t b_hi, uint32_t b_tap,
                                            uint32_t idesc, uint32_t first) {
  constexpr uint32_t ROW_UNITS = 2 * KSTEPS;   // row_bytes / 16
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const uint32_t a_lo = a_pl + (uint32_t)(kh * HALO_W + kw) * ROW_UNITS;
      const uint32_t b_lo = b_pl + (uint32_t)(kh * 3 + kw) * b_tap;
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k) {
        if (kh == 0 && kw == 0 && k == 0) umma_f16(tmem_d, desc64(a_lo, a_hi), desc64(b_lo, b_hi), idesc, first ? 0u : 1u);
        else umma_f16(tmem_d, desc64(a_lo + 2 * k, a_hi), desc64(b_lo + 2 * k, b_hi), idesc, 1u);
      }
    }
}

// `op` output planes at once: for every (kd, tap, k-step) one MMA per output plane o, so consecutive MMAs target DIFFERENT
// TMEM accumulators (and share the B descriptor).  a_st: first halo plane of the stage; output o reads plane o + kd.
template <int KSTEPS>
__device__ __forceinline__ void issue_planes_interleaved(uint32_t tmem0, uint32_t co, int op, uint32_t a_st, uint32_t plane_units, uint32_t a_hi,
                                                         uint32_t b_st, uint32_t b_plane, uint32_t b_hi, uint32_t b_tap, uint32_t idesc,
                                                         uint32_t first, uint32_t tap_on) {
  constexpr uint32_t ROW_UNITS = 2 * KSTEPS;
#pragma unroll
  for (int kd = 0; kd < 3; ++kd)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) {
          const uint32_t a_off = (uint32_t)kd * plane_units + tap_on * (uint32_t)(kh * HALO_W + kw) * ROW_UNITS + 2 * k;
          const uint32_t b_lo = b_st + (uint32_t)kd * b_plane + (uint32_t)(kh * 3 + kw) * b_tap + 2 * k;
          const uint32_t accf = (kd == 0 && kh == 0 && kw == 0 && k == 0) ? (first ? 0u : 1u) : 1u;
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < op) umma_f16(tmem0 + (uint32_t)o * co, desc64(a_st + (uint32_t)o * plane_units + a_off, a_hi), desc64(b_lo, b_hi), idesc, accf);
        }
}

// KSTEPS = kc/16; CO_T = 16 / 32: output channels known at compile time (register-resident statistics), 0 = generic
template <int KSTEPS, int CO_T>
__global__ void __launch_bounds__(192, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x2,
                 const __grid_constant__ CUtensorMap map_w2, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int wtiles = (27 + (p.mode2 ? 1 : 0)) * p.nchunk;
  const uint32_t wres_bytes = p.resident ? (uint32_t)wtiles * p.b_bytes : 0;
  uint8_t* ring = smem + wres_bytes;
  uint64_t* full = (uint64_t*)(ring + (size_t)p.stages * p.stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(wfull + 1);
  float* red = (float*)(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dbg = p.dbg && blockIdx.x == 0;
  trace_start(p.trace);
  if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();
  const int kd_groups = 3 / p.planes;                       // stages along kd per chunk
  const int nstage_per_tile = kd_groups * p.nchunk;         // (kd group, chunk)
  const uint32_t w_tx = (uint32_t)(p.Co * p.row_bytes);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.resident) {   // weight tile index = (tap * nchunk + chunk), tap = (kd*3+kh)*3+kw
      mbar_expect_tx(wfull, (uint32_t)wtiles * w_tx);
      for (int tap = 0; tap < 27; ++tap)
        for (int ch = 0; ch < p.nchunk; ++ch)
          tma_load_2d(smem_u32(smem) + (uint32_t)(tap * p.nchunk + ch) * p.b_bytes, &map_w, wfull, ch * p.kc, tap * p.Co);
      if (p.mode2)
        for (int ch = 0; ch < p.nchunk; ++ch)
          tma_load_2d(smem_u32(smem) + (uint32_t)(27 * p.nchunk + ch) * p.b_bytes, &map_w2, wfull, ch * p.kc, 0);
    }
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
    // ---- TMA producer: per (kd group, chunk) one halo box (+ its 9*planes weight tiles when not resident)
    int stage = 0; uint32_t phase = 0;
    const uint32_t ring_u = smem_u32(ring);
    const int box_planes = p.planes == 3 ? p.op + 2 : 1;
    const uint32_t tx = (uint32_t)(box_planes * p.plane_bytes) + (p.resident ? 0u : 9u * (uint32_t)p.planes * w_tx);
    int t = blockIdx.x;
    int n = t / p.tiles_per_n, r = t - n * p.tiles_per_n;
    for (; t < p.total_tiles; t += gridDim.x) {
      const int tw = r % p.tiles_w, q = r / p.tiles_w, th = q % p.tiles_h, d = (q / p.tiles_h) * p.op;   // first output plane of the tile
      for (int kg = 0; kg < kd_groups; ++kg)
        for (int ch = 0; ch < p.nchunk; ++ch) {
          mbar_wait(empty + stage, phase ^ 1);
          if (elect_one()) {
            const uint32_t base = ring_u + (uint32_t)stage * p.stage_bytes;
            if (p.dbg_mode & 1) mbar_arrive(full + stage);
            else {
              const bool second = p.mode2 == 2 && (p.planes == 3 || kg == 1);   // the stage that holds the centre plane
              mbar_expect_tx(full + stage, tx + (second ? (uint32_t)(p.op * 128 * p.row_bytes) : 0u));
              tma_load_5d(base, &map_x, full + stage, ch * p.kc, tw * HTW - 1, th * HTH - 1, d + kg * p.planes - 1, n);
              if (second) tma_load_5d(base + p.x2_off, &map_x2, full + stage, ch * p.kc, tw * HTW, th * HTH, d, n);
              if (!p.resident)
                for (int j = 0; j < 9 * p.planes; ++j)
                  tma_load_2d(base + p.halo_bytes + j * p.b_bytes, &map_w, full + stage, ch * p.kc, (kg * p.planes * 9 + j) * p.Co);
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      r += gridDim.x;
      while (r >= p.tiles_per_n) { r -= p.tiles_per_n; ++n; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Co >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t layout = KSTEPS == 4 ? 2u : (KSTEPS == 2 ? 4u : 6u);
    constexpr uint32_t ROW_BYTES = 32 * KSTEPS;
    const uint32_t a_hi = desc_hi((p.dbg_mode & 64) ? 8u * ROW_BYTES : (uint32_t)(HALO_W * ROW_BYTES), layout);   // next 8-row group = next h line of the halo
    const uint32_t b_hi = desc_hi(8 * ROW_BYTES, layout);
    const uint32_t ring_u = smem_u32(ring);
    const uint32_t a_lo0 = desc_lo(ring_u, 16);
    const uint32_t b_lo0 = p.resident ? desc_lo(smem_u32(smem), 16) : desc_lo(ring_u + p.halo_bytes, 16);
    const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4, b_units = (uint32_t)p.b_bytes >> 4, plane_units = (uint32_t)p.plane_bytes >> 4;
    const uint32_t b_tap = p.resident ? (uint32_t)p.nchunk * b_units : b_units;   // distance between consecutive taps' weight tiles
    const uint32_t b_plane = 9 * b_tap;                                           // ... and between consecutive kd planes
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    if (p.resident) { mbar_wait(wfull, 0); tc_fence_after(); }
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait(tempty + acc, acc_phase ^ 1);
      tc_fence_after();
      int s = 0;
      for (int kg = 0; kg < kd_groups; ++kg)
        for (int ch = 0; ch < p.nchunk; ++ch, ++s) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t a_st = a_lo0 + (uint32_t)stage * stage_units;
          const uint32_t b_st = p.resident ? b_lo0 + (uint32_t)(kg * p.planes) * b_plane + (uint32_t)ch * b_units : b_lo0 + (uint32_t)stage * stage_units;
          if (elect_one()) {
            if (!(p.dbg_mode & 2)) {
              const uint32_t b_k1 = b_lo0 + (uint32_t)(27 * p.nchunk + ch) * b_units;
              if (p.planes == 3) {
                // `op` output planes share the op+2 halo planes of this stage: output o reads planes o, o+1, o+2
                const bool inter = p.op > 1 && (p.dbg_mode & 32);   // measured slower than plane-by-plane (92 vs 87 us): opt-in experiment
                if (inter)
                  issue_planes_interleaved<KSTEPS>(tmem_base + (uint32_t)(acc * p.op * p.Co), (uint32_t)p.Co, p.op, a_st, plane_units, a_hi, b_st, b_plane,
                                                   b_hi, b_tap, idesc, s == 0 ? 1u : 0u, (p.dbg_mode & 16) ? 0u : 1u);
                for (int o = 0; o < p.op; ++o) {
                  const uint32_t tmem_d = tmem_base + (uint32_t)((acc * p.op + o) * p.Co);
                  const uint32_t a_o = a_st + (uint32_t)o * plane_units;
                  if (!inter) {
                    issue_plane<KSTEPS>(tmem_d, a_o, a_hi, b_st, b_hi, b_tap, idesc, s == 0 ? 1u : 0u);
                    issue_plane<KSTEPS>(tmem_d, a_o + plane_units, a_hi, b_st + b_plane, b_hi, b_tap, idesc, 0u);
                    issue_plane<KSTEPS>(tmem_d, a_o + 2 * plane_units, a_hi, b_st + 2 * b_plane, b_hi, b_tap, idesc, 0u);
                  }
                  if (p.mode2 == 1) {       // fused 1x1x1 conv: centre tap view of the centre plane, second accumulator set
                    const uint32_t a_c = a_o + plane_units + (uint32_t)(HALO_W + 1) * (2 * KSTEPS);
                    const uint32_t tmem_d2 = tmem_base + (uint32_t)(((2 + acc) * p.op + o) * p.Co);
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k)
                      umma_f16(tmem_d2, desc64(a_c + 2 * k, a_hi), desc64(b_k1 + 2 * k, b_hi), idesc, (ch | k) ? 1u : 0u);
                  } else if (p.mode2 == 2) { // fused 1x1x1 dgrad: second input tile (no halo), same accumulator
                    const uint32_t a_2 = a_st + ((uint32_t)p.x2_off >> 4) + (uint32_t)o * (128u * ROW_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k)
                      umma_f16(tmem_d, desc64(a_2 + 2 * k, b_hi), desc64(b_k1 + 2 * k, b_hi), idesc, 1u);
                  }
                }
              } else {
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.Co);
                issue_plane<KSTEPS>(tmem_d, a_st, a_hi, b_st, b_hi, b_tap, idesc, s == 0 ? 1u : 0u);
                if (p.mode2 && kg == 1) {   // the stage that holds the centre plane
                  if (p.mode2 == 1) {
                    const uint32_t a_c = a_st + (uint32_t)(HALO_W + 1) * (2 * KSTEPS);
                    const uint32_t tmem_d2 = tmem_base + (uint32_t)((2 + acc) * p.Co);
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k)
                      umma_f16(tmem_d2, desc64(a_c + 2 * k, a_hi), desc64(b_k1 + 2 * k, b_hi), idesc, (ch | k) ? 1u : 0u);
                  } else {
                    const uint32_t a_2 = a_st + ((uint32_t)p.x2_off >> 4);
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k)
                      umma_f16(tmem_d, desc64(a_2 + 2 * k, b_hi), desc64(b_k1 + 2 * k, b_hi), idesc, 1u);
                  }
                }
              }
            }
            umma_commit(empty + stage);
            if (s == nstage_per_tile - 1) umma_commit(tfull + acc);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ---- epilogue (TMEM lane quarter = warp % 4); row r -> (h = r/8, w = r%8) of the d-plane tile
    const int q4 = warp & 3, ew = warp - 2;
    const int row = q4 * 32 + lane;
    const int npass = p.mode2 == 1 ? 2 : 1;      // pass 1 = the fused 1x1x1 output (second accumulator pair)
    int acc = 0; uint32_t acc_phase = 0;
    int t = blockIdx.x;
    int n = t / p.tiles_per_n, r = t - n * p.tiles_per_n;
    constexpr int NR = CO_T ? CO_T : 1;
    float rs1[2][NR], rs2[2][NR];    // register-resident statistics (CO_T != 0), per pass
#pragma unroll
    for (int j = 0; j < NR; ++j) { rs1[0][j] = rs2[0][j] = rs1[1][j] = rs2[1][j] = 0.f; }
    int n_acc = n;
    auto flush = [&](int nn) {       // all 128 epilogue threads
      if (CO_T) {
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
          double* sp = ps ? p.stats2 : p.stats;
          if (ps >= npass || !sp) continue;
#pragma unroll
          for (int j = 0; j < NR; ++j) {
            float s1 = warp_sum(rs1[ps][j]), s2 = warp_sum(rs2[ps][j]);
            if (lane == 0) { red[ew * 2 * CO_T + j] = s1; red[ew * 2 * CO_T + CO_T + j] = s2; }
            rs1[ps][j] = 0.f; rs2[ps][j] = 0.f;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          const int e = ew * 32 + lane;
          if (e < 2 * CO_T) {
            float tot = red[e] + red[2 * CO_T + e] + red[4 * CO_T + e] + red[6 * CO_T + e];
            int c = e % NR, which = e / NR;
            atomicAdd(sp + ((long)nn * CO_T + c) * 2 + which, (double)tot);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
    };
    for (; t < p.total_tiles; t += gridDim.x) {
      const int tw = r % p.tiles_w, qq = r / p.tiles_w, th = qq % p.tiles_h, d0 = (qq / p.tiles_h) * p.op;
      if (n != n_acc) { flush(n_acc); n_acc = n; }
      const int w = tw * HTW + (row & 7), h = th * HTH + (row >> 3);
      const bool valid = (w < p.W) && (h < p.H);
      mbar_wait(tfull + acc, acc_phase);
      tc_fence_after();
      for (int o = 0; o < p.op; ++o) {
      const int d = d0 + o;
      if (d >= p.D) break;
      const long vox = (((long)n * p.D + d) * p.H + h) * p.W + w;
#pragma unroll
      for (int ps = 0; ps < 2; ++ps) {
        if (ps >= npass) continue;
        bf16* dst = ps ? p.out2 + vox * p.pitch2 + p.coff2 : p.out + vox * p.pitch + p.coff;
        double* sp = ps ? p.stats2 : p.stats;
        const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(((2 * ps + acc) * p.op + o) * p.Co);
        if (CO_T) {
#pragma unroll
          for (int c0 = 0; c0 < NR; c0 += 16) {
            float v[16];
            tmem_ld16(trow + c0, v);
            if (p.dbg_mode & 4) continue;
            if (valid) {
              if (sp) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { rs1[ps][(c0 + j) % NR] += v[j]; rs2[ps][(c0 + j) % NR] = fmaf(v[j], v[j], rs2[ps][(c0 + j) % NR]); }
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
        } else {
          for (int c0 = 0; c0 < p.Co; c0 += 16) {
            float v[16];
            tmem_ld16(trow + c0, v);
            if (p.dbg_mode & 4) continue;
            if (sp) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float x = valid ? v[j] : 0.f;
                float s1 = warp_sum(x), s2 = warp_sum(x * x);
                if (lane == 0) { red[ew * 2 * p.Co + c0 + j] = s1; red[ew * 2 * p.Co + p.Co + c0 + j] = s2; }
              }
            }
            if (valid) {
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
          if (sp) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            int e = ew * 32 + lane;
            for (int i = e; i < 2 * p.Co; i += 128) {
              float tot = red[i] + red[2 * p.Co + i] + red[4 * p.Co + i] + red[6 * p.Co + i];
              int c = i % p.Co, which = i / p.Co;
              atomicAdd(sp + ((long)n * p.Co + c) * 2 + which, (double)tot);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
          }
        }
      }
      }   // o
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
  if (dbg && threadIdx.x == 0) p.dbg[48] = clock64();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// smem plan shared by the support test and the launcher
struct HaloPlan { int kc, rb, nchunk, plane_bytes, b_bytes, resident, planes, op, halo_bytes, x2_bytes, stage_bytes, stages; };
// D, tiles_hw: depth and (h,w) tile count of one sample -- the multi-plane tiling needs enough tiles to balance 148 CTAs
static inline HaloPlan halo_plan(int Ci, int Co, int mode2 = 0, int D = 0, long tiles_hw_n = 0) {
  HaloPlan h;
  h.kc = Ci % 64 == 0 ? 64 : (Ci % 32 == 0 ? 32 : 16); h.rb = h.kc * 2; h.nchunk = Ci / h.kc;
  h.plane_bytes = HALO_H * HALO_W * h.rb;
  h.b_bytes = ((Co * h.rb + 1023) / 1024) * 1024;
  const int budget_all = 200 * 1024;
  const int wtiles = (27 + (mode2 ? 1 : 0)) * h.nchunk;
  h.resident = ((long)wtiles * h.b_bytes <= 112 * 1024) ? 1 : 0;
  int budget = budget_all - (h.resident ? wtiles * h.b_bytes : 0);
  const int x2_one = mode2 == 2 ? 128 * h.rb : 0;   // second input tile of the fused 1x1x1 dgrad, per output plane
  const int op_max = getenv("B200_HALO_OP") ? atoi(getenv("B200_HALO_OP")) : 4;
  const long min_tiles = getenv("B200_HALO_MIN_TILES") ? atol(getenv("B200_HALO_MIN_TILES")) : 148 * 8;   // tests lower it
  h.planes = 3; h.op = 1;
  // most output planes per tile that keep >= 3 pipeline stages, fit TMEM (2 buffers x op x Co [x2 outputs]) and leave every CTA
  // at least ~8 tiles; resident weights and a single channel chunk only
  for (int op = 4; op >= 1; op >>= 1) {
    if (op > op_max) continue;
    if (op > 1 && (!h.resident || h.nchunk != 1 || D <= 0 || D % op != 0 || (long)(D / op) * tiles_hw_n < min_tiles)) continue;
    if ((mode2 == 1 ? 4 : 2) * op * Co > 512) continue;
    const int hb = (((op + 2) * h.plane_bytes + 1023) / 1024) * 1024, xb = ((op * x2_one + 1023) / 1024) * 1024;
    const int st = hb + xb + (h.resident ? 0 : 27 * h.b_bytes);
    if (budget / st < 3 && op > 1) continue;
    h.op = op; h.halo_bytes = hb; h.x2_bytes = xb; h.stage_bytes = st;
    break;
  }
  if (budget / h.stage_bytes < 3) {   // one plane per stage
    h.planes = 1; h.op = 1;
    h.halo_bytes = ((h.plane_bytes + 1023) / 1024) * 1024;
    h.x2_bytes = ((x2_one + 1023) / 1024) * 1024;
    h.stage_bytes = h.halo_bytes + h.x2_bytes + (h.resident ? 0 : 9 * h.b_bytes);
  }
  h.stages = budget / h.stage_bytes; if (h.stages > 8) h.stages = 8;
  return h;
}
static inline bool conv_halo_supported(int Ci, int Co) {
  if (!(Ci % 16 == 0 && Co % 16 == 0 && Co <= 256)) return false;
  return halo_plan(Ci, Co).stages >= 2;
}
// fused 3x3x3 + 1x1x1 (mode2 1: two outputs of one input, 2: two inputs of one output): resident weights, 4*Co TMEM columns
static inline bool conv_halo_fused_supported(int Ci, int Co, int mode2) {
  if (!(Ci % 16 == 0 && Co % 16 == 0) || getenv("B200_NO_FUSED_K1")) return false;
  if (mode2 == 1 && 4 * Co > 512) return false;
  HaloPlan h = halo_plan(Ci, Co, mode2);
  return h.resident && h.stages >= 2;
}

template <int KSTEPS, int CO_T>
static int conv_halo_launch(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& mx2, const CUtensorMap& mw2, const HaloParams& p, int grid,
                            size_t smem, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) { B200_CUDA(cudaFuncSetAttribute(conv_halo_kernel<KSTEPS, CO_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_done = true; }
  B200_CUDA(launch_pdl(conv_halo_kernel<KSTEPS, CO_T>, dim3(grid), dim3(192), smem, st, mx, mw, mx2, mw2, p));
  B200_LAUNCH_CHECK();
  return 0;
}

// 3x3x3 only.  wp: packed bf16 [27][Co][Ci] (same packing as tc::conv).
// optional (dgrad of a conv whose input was a = lrelu(norm(c))): the first pass of that norm's backward rides on the epilogue -- `acc`
// ([N][C][3] doubles, zeroed) receives (sum g, sum g*n); *done tells the caller whether the kernel that ran supports it
struct HaloNormBwd { const bf16* act; int pitch, coff; double* acc; bool* done; };
struct HaloFused {   // optional fused 1x1x1 conv (see HaloParams::mode2); wp2: packed bf16 [Co][Ci]
  int mode2; const bf16* wp2;
  bf16* out2; int pitch2, coff2; double* stats2;          // mode2 == 1
  const bf16* x2; int x2_pitch, x2_coff;                   // mode2 == 2 (same channel count Ci as x)
};
static int conv_halo(const bf16* x, int in_pitch, int in_coff, int Ci, int N, int D, int H, int W, const bf16* wp, int Co,
                     bf16* out, int out_pitch, int out_coff, int accumulate, double* stats, cudaStream_t st, const HaloFused* fu = nullptr,
                     int out_half = 0, const HaloNormBwd* nb = nullptr) {
  if (nb && nb->done) *nb->done = false;
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
  const int mode2 = fu ? fu->mode2 : 0;
  {  // 16 / 32 output channels: kw taps stacked along N (tc_conv_halo48.cuh), 9 instead of 27 MMAs per k-step
    Halo48Plan h48 = halo48_plan(Ci, Co, mode2, D, (long)N * cdiv(H, S48_TH) * cdiv(W, S48_TW));
    if (h48.ok) {
      const bool fold = nb && !stats && !mode2 && !accumulate && (nb->pitch % 8 == 0) && (nb->coff % 8 == 0);
      if (fold && nb->done) *nb->done = true;
      return conv_halo48(h48, x, in_pitch, in_coff, Ci, N, D, H, W, wp, Co, out, out_pitch, out_coff, accumulate, fold ? nb->acc : stats, st, mode2,
                         fu ? fu->wp2 : nullptr, fu ? fu->out2 : nullptr, fu ? fu->pitch2 : 0, fu ? fu->coff2 : 0, fu ? fu->stats2 : nullptr,
                         fu ? fu->x2 : nullptr, fu ? fu->x2_pitch : 0, fu ? fu->x2_coff : 0, out_half, fold ? nb->act : nullptr, fold ? nb->pitch : 0,
                         fold ? nb->coff : 0);
    }
  }
  HaloPlan h = halo_plan(Ci, Co, mode2, D, (long)N * cdiv(H, HTH) * cdiv(W, HTW));
  B200_CHECK(!mode2 || h.resident, "fused 1x1x1 conv needs resident weights (Ci=%d Co=%d)", Ci, Co);
  HaloParams p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co;
  p.kc = h.kc; p.row_bytes = h.rb; p.nchunk = h.nchunk;
  p.tiles_w = cdiv(W, HTW); p.tiles_h = cdiv(H, HTH);
  p.op = h.op;
  p.tiles_per_n = cdiv(D, h.op) * p.tiles_h * p.tiles_w;
  long total = (long)N * p.tiles_per_n;
  B200_CHECK(total < (1L << 30), "halo conv: too many tiles");
  p.total_tiles = (int)total;
  p.planes = h.planes; p.plane_bytes = h.plane_bytes; p.halo_bytes = h.halo_bytes; p.b_bytes = h.b_bytes; p.resident = h.resident;
  p.stage_bytes = h.stage_bytes; p.stages = h.stages;
  B200_CHECK(p.stages >= 2, "halo conv smem budget exceeded (Ci=%d Co=%d)", Ci, Co);
  uint32_t cols = (mode2 == 1 ? 4 : 2) * h.op * Co, pw = 32; while (pw < cols) pw <<= 1; p.tmem_cols = pw;
  B200_CHECK(p.tmem_cols <= 512, "halo conv TMEM budget exceeded");
  p.out = out; p.pitch = out_pitch; p.coff = out_coff; p.accumulate = accumulate; p.stats = stats;
  p.out_half = out_half;
  p.mode2 = mode2; p.out2 = nullptr; p.pitch2 = p.coff2 = 0; p.stats2 = nullptr; p.x2_off = h.halo_bytes;
  if (mode2 == 1) { p.out2 = fu->out2; p.pitch2 = fu->pitch2; p.coff2 = fu->coff2; p.stats2 = fu->stats2; }
  p.dbg = g_dbg; p.dbg_mode = 0;
  p.trace = trace_slot(); if (p.trace) trace_tag("conv_halo %d->%d @%d mode2=%d", Ci, Co, D, mode2);
  if (const char* e = getenv("B200_HALO_DBG")) p.dbg_mode = atoi(e);
  if (const char* e = getenv("B200_HALO_STAGES")) { int v = atoi(e); if (v >= 2 && v <= p.stages) p.stages = v; }

  CUtensorMapSwizzle sw = p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap mx, mw, mx2, mw2;
  memset(&mx2, 0, sizeof(mx2)); memset(&mw2, 0, sizeof(mw2));
  if (mode2 == 2) {
    cuuint64_t dims[5] = {(cuuint64_t)Ci, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    const int pt = fu->x2_pitch;
    cuuint64_t strides[4] = {(cuuint64_t)pt * 2, (cuuint64_t)W * pt * 2, (cuuint64_t)H * W * pt * 2, (cuuint64_t)D * H * W * pt * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.kc, HTW, HTH, (cuuint32_t)h.op, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&mx2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(fu->x2 + fu->x2_coff), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "halo conv second-input tensor map failed (%d)", (int)r);
  }
  if (mode2) {
    cuuint64_t dims[2] = {(cuuint64_t)Ci, (cuuint64_t)Co};
    cuuint64_t strides[1] = {(cuuint64_t)Ci * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)Co};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mw2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)fu->wp2, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "halo conv 1x1x1 weight tensor map failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)Ci, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)W * in_pitch * 2, (cuuint64_t)H * W * in_pitch * 2, (cuuint64_t)D * H * W * in_pitch * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.kc, HALO_W, HALO_H, (cuuint32_t)(p.planes == 3 ? h.op + 2 : 1), 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(x + in_coff), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "halo conv input tensor map failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Ci, (cuuint64_t)27 * Co};
    cuuint64_t strides[1] = {(cuuint64_t)Ci * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)Co};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)wp, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "halo conv weight tensor map failed (%d)", (int)r);
  }
  size_t smem = (size_t)p.stages * p.stage_bytes + (p.resident ? (size_t)(27 + (mode2 ? 1 : 0)) * p.nchunk * p.b_bytes : 0) + 1024 + 256 + 8 * Co * sizeof(float) + 64;
  B200_CHECK(smem <= 227 * 1024, "halo conv smem budget exceeded (%zu)", smem);
  int grid = (int)(p.total_tiles < num_sms() ? p.total_tiles : num_sms());
  const int ks = p.kc / 16;
#define B200_HALO_CASE(KS, CO) if (ks == KS && ((CO) ? Co == (CO) : (Co != 16 && Co != 32))) return conv_halo_launch<KS, CO>(mx, mw, mx2, mw2, p, grid, smem, st)
  B200_HALO_CASE(1, 16); B200_HALO_CASE(1, 32); B200_HALO_CASE(1, 0);
  B200_HALO_CASE(2, 16); B200_HALO_CASE(2, 32); B200_HALO_CASE(2, 0);
  B200_HALO_CASE(4, 16); B200_HALO_CASE(4, 32); B200_HALO_CASE(4, 0);
#undef B200_HALO_CASE
  B200_CHECK(false, "halo conv: no kernel for kc=%d Co=%d", p.kc, Co);
}

}  // namespace tc
}  // namespace b200
