// tcgen05 implicit-GEMM 3-D convolution (k in {1,3}, stride 1, "same" zero padding, bias-free) on channels-last bf16.
//
//   out[v, co] = sum_{tap, ci} x[v + tap - pad, ci] * Wp[tap][co][ci]
//
// GEMM view: M = 128 voxels (a 4x4x8 brick of the volume), N = Co, K = taps * Ci.  No im2col buffer exists anywhere:
// for every tap the TMA engine fetches the brick shifted by (kd-1,kh-1,kw-1) straight from the 5-D tensor
// (C, W, H, D, N) into a K-major swizzled smem tile; out-of-volume voxels are zero-filled by TMA, which *is* the
// padding.  The same kernel computes dgrad (input = dy, weights packed flipped/transposed by pack_conv_weights).
// Warp roles as in tc_gemm.cuh.  Epilogue: TMEM -> registers -> bf16 channels-last rows (16-byte stores), optionally
// accumulating into the destination (dx = dgrad3x3 + dgrad1x1) and emitting per-(n,c) sum / sum-of-squares of the
// fp32 accumulators for the InstanceNorm that follows (K10 -> K12 fusion of SURVEY 2.1).
#pragma once
#include "tc_gemm.cuh"

namespace b200 {
namespace tc {

static constexpr int CONV_PRODUCERS = 1;   // TMA-issuing warps (measured: more than one does not help, the TMA unit is row-bound)
static constexpr int CONV_LOADERS = 128;   // threads of the software A-brick loader (warps 6..9)
static constexpr int CONV_THREADS = 192 + CONV_LOADERS;
static constexpr int TD = 4, TH = 4, TW = 8;  // voxel brick = 128 GEMM rows (w fastest: matches the TMA box order)

struct ConvParams {
  int N, D, H, W, Ci, Co, ks;
  int kc, row_bytes;          // channels per k-block and its smem row size (32/64/128 B -> swizzle mode)
  int nchunk, kblocks;        // Ci/kc ; taps*nchunk
  int group;                  // k-blocks per pipeline stage
  int tiles_w, tiles_h, tiles_d; long total_tiles;
  int a_bytes, b_bytes;       // per k-block, b rounded up to 1024
  int stages; uint32_t tmem_cols;
  // epilogue
  bf16* out; int pitch, coff, accumulate; double* stats; int out_half;
  long long* dbg;   // optional CTA-0 clock64 stamps (tuning aid)
  long long* trace;
  int resident;     // all packed weights (kblocks x b_bytes) stay in smem for the CTA's lifetime; stages hold A bricks only
  // Software A loader (row_bytes <= 64): TMA needs ~2 cycles per 32-byte box row, so the 27 shifted bricks are instead
  // fetched with coalesced 16-byte ld.global (the 27x overlap hits L1), written to smem in the hardware swizzle pattern,
  // and published to the tensor core with fence.proxy.async + mbarrier.
  int sw_loader; const bf16* x; int in_pitch;
};

__device__ __forceinline__ uint64_t smem_desc_k(uint32_t addr, int row_bytes) {
  // K-major canonical layout, 8-row groups `8*row_bytes` apart; layout_type: 128B=2, 64B=4, 32B=6
  uint64_t lt = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(((uint32_t)(8 * row_bytes) >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= lt << 61;
  return d;
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

static __global__ void __launch_bounds__(CONV_THREADS, 1)
conv_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t kb_bytes = p.resident ? p.a_bytes : p.a_bytes + p.b_bytes, stage_bytes = kb_bytes * p.group;
  const uint32_t wres_bytes = p.resident ? (uint32_t)p.kblocks * p.b_bytes : 0;   // resident weights live in front of the ring
  uint8_t* ring = smem + wres_bytes;
  uint64_t* full = (uint64_t*)(ring + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(wfull + 1);
  float* red = (float*)(tmem_slot + 4);  // [4 warps][2*Co] partial statistics

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dbg = p.dbg && blockIdx.x == 0;
  trace_start(p.trace);
  if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();
  const int nstage_per_tile = (p.kblocks + p.group - 1) / p.group;
  const int pad = p.ks >> 1;

  if (threadIdx.x == 0) {
    const uint32_t full_count = p.sw_loader ? (uint32_t)CONV_LOADERS + (p.resident ? 0u : 1u) : 1u;
    for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, full_count); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.resident) {   // fetch every (tap, chunk) weight tile once
      mbar_expect_tx(wfull, (uint32_t)p.kblocks * (uint32_t)(p.Co * p.row_bytes));
      int tap = 0, ch = 0;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        tma_load_2d(smem_u32(smem) + (uint32_t)kb * p.b_bytes, &map_w, wfull, ch * p.kc, tap * p.Co);
        if (++ch == p.nchunk) { ch = 0; ++tap; }
      }
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
  pdl_wait();   // (resident weights fetched above were packed many launches earlier)

  if (warp >= 6) {
    // ---- software A-brick loader (CONV_LOADERS threads)
    if (p.sw_loader) {
      const int lt = threadIdx.x - 192;
      const int cpr = p.row_bytes >> 4;                 // 16-byte chunks per row: 2 (32 B) or 4 (64 B) = chunks per thread
      const uint32_t ring_u = smem_u32(ring);
      const long sW = p.in_pitch, sH = (long)p.W * p.in_pitch, sD = (long)p.H * p.W * p.in_pitch;   // element strides
      int stage = 0; uint32_t phase = 0;
      for (long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        int tw = (int)(t % p.tiles_w); long r = t / p.tiles_w; int th = (int)(r % p.tiles_h); r /= p.tiles_h;
        int td = (int)(r % p.tiles_d); int n = (int)(r / p.tiles_d);
        // tap-invariant per-chunk state: q = lt + j*128 -> row = q / cpr, c = q % cpr (consecutive threads = consecutive bytes)
        const bf16* bp[4]; uint32_t so[4]; uint32_t mk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int q = lt + j * CONV_LOADERS, row = q / cpr, c = q - row * cpr;
          int w = tw * TW + (row & 7), h = th * TH + ((row >> 3) & 3), d = td * TD + (row >> 5);
          bp[j] = p.x + ((((long)n * p.D + d) * p.H + h) * p.W + w) * p.in_pitch + c * 8;
          int cs = cpr == 2 ? (c ^ ((row >> 2) & 1)) : (c ^ ((row >> 1) & 3));   // hardware 32 B / 64 B swizzle
          so[j] = (uint32_t)row * p.row_bytes + (uint32_t)cs * 16;
          uint32_t m = 0;   // bit k: coordinate + k - pad inside the volume, for w (bits 0-2), h (3-5), d (6-8)
          for (int k = 0; k < p.ks; ++k) {
            m |= ((unsigned)(w + k - pad) < (unsigned)p.W ? 1u : 0u) << k;
            m |= ((unsigned)(h + k - pad) < (unsigned)p.H ? 1u : 0u) << (3 + k);
            m |= ((unsigned)(d + k - pad) < (unsigned)p.D ? 1u : 0u) << (6 + k);
          }
          mk[j] = m;
        }
        int kw = 0, kh = 0, kd = 0, ch = 0;
        for (int s = 0; s < nstage_per_tile; ++s) {
          const int cnt = min(p.group, p.kblocks - s * p.group);
          mbar_wait(empty + stage, phase ^ 1);
          const uint32_t base = ring_u + (uint32_t)stage * stage_bytes;
          for (int i0 = 0; i0 < cnt; i0 += 4) {           // sub-batches of <= 4 k-blocks: all loads first, then the stores
            uint4 v[4][4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              if (i0 + ii < cnt) {
                const long delta = (kd - pad) * sD + (kh - pad) * sH + (kw - pad) * sW + ch * p.kc;   // warp-uniform
                const uint32_t need = (1u << kw) | (8u << kh) | (64u << kd);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (j < cpr) v[ii][j] = ((mk[j] & need) == need) ? *reinterpret_cast<const uint4*>(bp[j] + delta) : make_uint4(0, 0, 0, 0);
                if (++ch == p.nchunk) { ch = 0; if (++kw == p.ks) { kw = 0; if (++kh == p.ks) { kh = 0; ++kd; } } }
              }
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              if (i0 + ii < cnt) {
                const uint32_t dst0 = base + (uint32_t)(i0 + ii) * kb_bytes;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  if (j < cpr)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst0 + so[j]), "r"(v[ii][j].x), "r"(v[ii][j].y), "r"(v[ii][j].z), "r"(v[ii][j].w) : "memory");
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
          mbar_arrive(full + stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 0) {
    // ---- TMA producer: warp-uniform loop, one elected lane issues.  With the software A loader it only streams the
    // weight tiles (when they are not resident); it arms the barrier with arrive + expect_tx for the bytes TMA delivers.
    const int pid = 0;
    int stage = 0; uint32_t phase = 0;
    const uint32_t smem_base_u = smem_u32(ring);
    const uint32_t kb_tx = (p.sw_loader ? 0u : (uint32_t)p.a_bytes) + (p.resident ? 0u : (uint32_t)(p.Co * p.row_bytes));   // bytes TMA delivers per k-block
    const bool idle = p.sw_loader && p.resident;   // nothing for TMA to do in the main loop
    for (long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int tw = (int)(t % p.tiles_w); long r = t / p.tiles_w; int th = (int)(r % p.tiles_h); r /= p.tiles_h;
      int td = (int)(r % p.tiles_d); int n = (int)(r / p.tiles_d);
      const int w0 = tw * TW - pad, h0 = th * TH - pad, d0 = td * TD - pad;
      int kw = 0, kh = 0, kd = 0, ch = 0, wrow = 0;   // running tap / chunk state (kb = tap*nchunk + ch)
      for (int s = 0; s < nstage_per_tile && !idle; ++s) {
        const int cnt = min(p.group, p.kblocks - s * p.group);
        mbar_wait(empty + stage, phase ^ 1);
        const uint32_t base = smem_base_u + (uint32_t)stage * stage_bytes;
        const bool leader = elect_one();
        if (leader && pid == 0) mbar_expect_tx(full + stage, (uint32_t)cnt * kb_tx);
        for (int i = 0; i < cnt; ++i) {
          if (leader && (i % CONV_PRODUCERS) == pid) {
            const uint32_t sa = base + i * kb_bytes;
            if (!p.sw_loader) tma_load_5d(sa, &map_x, full + stage, ch * p.kc, w0 + kw, h0 + kh, d0 + kd, n);
            if (!p.resident) tma_load_2d(sa + p.a_bytes, &map_w, full + stage, ch * p.kc, wrow);
          }
          if (++ch == p.nchunk) { ch = 0; wrow += p.Co; if (++kw == p.ks) { kw = 0; if (++kh == p.ks) { kh = 0; ++kd; } } }
        }
        if (dbg && leader && pid == 0 && t < blockIdx.x + 2 * (long)gridDim.x) p.dbg[1 + (t != blockIdx.x) * 8 + s] = clock64();
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Co >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t layout = p.row_bytes == 128 ? 2u : (p.row_bytes == 64 ? 4u : 6u);
    const uint32_t hi = desc_hi(8 * p.row_bytes, layout);
    const uint32_t smem_base_u = smem_u32(ring);
    const uint32_t a_lo0 = desc_lo(smem_base_u, 16), b_lo0 = p.resident ? desc_lo(smem_u32(smem), 16) : desc_lo(smem_base_u + p.a_bytes, 16);
    const uint32_t stage_units = stage_bytes >> 4, kb_units = kb_bytes >> 4, b_units = (p.resident ? (uint32_t)p.b_bytes : kb_bytes) >> 4;
    if (p.resident) { mbar_wait(wfull, 0); tc_fence_after(); }
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    const int ksteps = p.kc / 16;
    for (long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait(tempty + acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.Co);
      for (int s = 0; s < nstage_per_tile; ++s) {
        const int cnt = min(p.group, p.kblocks - s * p.group);
        mbar_wait(full + stage, phase);
        tc_fence_after();
        const uint32_t a_st = a_lo0 + (uint32_t)stage * stage_units;
        const uint32_t b_st = p.resident ? b_lo0 + (uint32_t)(s * p.group) * b_units : b_lo0 + (uint32_t)stage * stage_units;
        if (dbg && lane == 0 && t < blockIdx.x + 2 * (long)gridDim.x) p.dbg[17 + (t != blockIdx.x) * 8 + s] = clock64();
        if (elect_one()) {
          uint32_t a_lo = a_st, b_lo = b_st;
          for (int i = 0; i < cnt; ++i) {
            for (int k = 0; k < ksteps; ++k)
              umma_f16(tmem_d, desc64(a_lo + 2 * k, hi), desc64(b_lo + 2 * k, hi), idesc, (s | i | k) ? 1u : 0u);
            a_lo += kb_units; b_lo += b_units;
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
    const int q = warp & 3, ew = warp - 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int tw = (int)(t % p.tiles_w); long r = t / p.tiles_w; int th = (int)(r % p.tiles_h); r /= p.tiles_h;
      int td = (int)(r % p.tiles_d); int n = (int)(r / p.tiles_d);
      const int row = q * 32 + lane;
      const int w = tw * TW + (row & 7), h = th * TH + ((row >> 3) & 3), d = td * TD + (row >> 5);
      const bool valid = (w < p.W) && (h < p.H) && (d < p.D);
      bf16* dst = p.out + ((((long)n * p.D + d) * p.H + h) * p.W + w) * p.pitch + p.coff;
      mbar_wait(tfull + acc, acc_phase);
      tc_fence_after();
      { long ti = (t - blockIdx.x) / gridDim.x; if (dbg && warp == 2 && lane == 0 && ti < 4) p.dbg[33 + 2 * ti] = clock64(); }
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Co);
      for (int c0 = 0; c0 < p.Co; c0 += 16) {
        float v[16];
        tmem_ld16(trow + c0, v);
        if (p.stats) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float x = valid ? v[j] : 0.f;
            float s1 = warp_sum(x), s2 = warp_sum(x * x);
            if (lane == 0) { red[ew * 2 * p.Co + c0 + j] = s1; red[ew * 2 * p.Co + p.Co + c0 + j] = s2; }
          }
        }
        if (valid) {
          if (p.accumulate) {
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
      if (p.stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        int e = (warp - 2) * 32 + lane;  // 0..127
        for (int i = e; i < 2 * p.Co; i += 128) {
          float tot = red[i] + red[2 * p.Co + i] + red[4 * p.Co + i] + red[6 * p.Co + i];
          int c = i % p.Co, which = i / p.Co;
          atomicAdd(p.stats + ((long)n * p.Co + c) * 2 + which, (double)tot);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      { long ti = (t - blockIdx.x) / gridDim.x; if (dbg && warp == 2 && lane == 0 && ti < 4) p.dbg[34 + 2 * ti] = clock64(); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
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

static inline bool conv_supported(int Ci, int Co, int in_pitch, int in_coff, int out_pitch, int out_coff) {
  return Ci % 16 == 0 && Co % 16 == 0 && Co <= 256 && in_pitch % 8 == 0 && in_coff % 8 == 0 && out_pitch % 8 == 0 && out_coff % 8 == 0;
}

// x: channels-last window (p + coff, pitch) with Ci channels; wp: packed bf16 [taps][Co][Ci]
static int conv(const bf16* x, int in_pitch, int in_coff, int Ci, int N, int D, int H, int W, const bf16* wp, int Co, int ks,
                bf16* out, int out_pitch, int out_coff, int accumulate, double* stats, cudaStream_t st, int out_half = 0) {
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
  ConvParams p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.ks = ks;
  p.kc = Ci % 64 == 0 ? 64 : (Ci % 32 == 0 ? 32 : 16);
  p.row_bytes = p.kc * 2;
  p.nchunk = Ci / p.kc; p.kblocks = ks * ks * ks * p.nchunk;
  p.a_bytes = 128 * p.row_bytes; p.b_bytes = ((Co * p.row_bytes + 1023) / 1024) * 1024;
  // measured on B200: the ld.global loader is 2x SLOWER than TMA here (both are bound by L2->SM request bandwidth, and the
  // small L1 left beside 200 KB of smem does not catch the 27x overlap) -- kept as an opt-in experiment only.
  p.sw_loader = 0;
  if (const char* e = getenv("B200_CONV_SWLOADER")) p.sw_loader = (atoi(e) && p.row_bytes <= 64 && in_coff % 8 == 0 && in_pitch % 8 == 0) ? 1 : 0;
  p.x = x + in_coff; p.in_pitch = in_pitch;
  p.resident = ((long)p.kblocks * p.b_bytes <= 64 * 1024) ? 1 : 0;
  if (const char* e = getenv("B200_CONV_RESIDENT")) p.resident = p.resident && atoi(e);
  int kb_bytes = p.resident ? p.a_bytes : p.a_bytes + p.b_bytes;
  int ring_budget = 196 * 1024 - (p.resident ? p.kblocks * p.b_bytes : 0);
  p.group = 32 * 1024 / kb_bytes; if (p.group < 1) p.group = 1; if (p.group > 9) p.group = 9; if (p.group > p.kblocks) p.group = p.kblocks;
  p.stages = (int)(ring_budget / (kb_bytes * p.group)); if (p.stages > 8) p.stages = 8; if (p.stages < 2) p.stages = 2;
  p.tiles_w = cdiv(W, TW); p.tiles_h = cdiv(H, TH); p.tiles_d = cdiv(D, TD);
  p.total_tiles = (long)N * p.tiles_d * p.tiles_h * p.tiles_w;
  uint32_t cols = 2 * Co, pw = 32; while (pw < cols) pw <<= 1; p.tmem_cols = pw;
  p.dbg = g_dbg;
  p.trace = trace_slot(); if (p.trace) trace_tag("conv k%d %d->%d @%d", ks, Ci, Co, D);
  p.out = out; p.pitch = out_pitch; p.coff = out_coff; p.accumulate = accumulate; p.stats = stats; p.out_half = out_half;

  CUtensorMapSwizzle sw = p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap mx, mw;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Ci, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)W * in_pitch * 2, (cuuint64_t)H * W * in_pitch * 2, (cuuint64_t)D * H * W * in_pitch * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.kc, TW, TH, TD, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(x + in_coff), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "conv input tensor map failed (%d): C=%d pitch=%d dims %dx%dx%d", (int)r, Ci, in_pitch, D, H, W);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Ci, (cuuint64_t)ks * ks * ks * Co};
    cuuint64_t strides[1] = {(cuuint64_t)Ci * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)Co};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)wp, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "conv weight tensor map failed (%d)", (int)r);
  }
  size_t smem = (size_t)p.stages * kb_bytes * p.group + (p.resident ? (size_t)p.kblocks * p.b_bytes : 0) + 1024 + 256 + 8 * Co * sizeof(float) + 64;
  B200_CHECK(smem <= 227 * 1024, "conv smem budget exceeded (%zu)", smem);
  static bool attr_done = false;
  if (!attr_done) { B200_CUDA(cudaFuncSetAttribute(conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_done = true; }
  int grid = (int)(p.total_tiles < num_sms() ? p.total_tiles : num_sms());
  B200_CUDA(launch_pdl(conv_kernel, dim3(grid), dim3(CONV_THREADS), smem, st, mx, mw, p));
  B200_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc

// W fp32 [Co][Ci][taps] -> fwd[tap][co][ci] and dgr[taps-1-tap][ci][co] (bf16)
static __global__ void pack_conv_weights_kernel(const float* __restrict__ W, bf16* __restrict__ fwd, bf16* __restrict__ dgr, int Co, int Ci, int taps) {
  long total = (long)Co * Ci * taps;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int tap = (int)(e % taps); long r = e / taps; int ci = (int)(r % Ci); int co = (int)(r / Ci);
    bf16 v = __float2bfloat16_rn(W[e]);
    fwd[((long)tap * Co + co) * Ci + ci] = v;
    dgr[((long)(taps - 1 - tap) * Ci + ci) * Co + co] = v;
  }
}

}  // namespace b200
