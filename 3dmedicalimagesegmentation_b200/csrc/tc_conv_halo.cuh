// tcgen05 3x3x3 convolution with on-SM halo reuse (forward and dgrad), channels-last bf16.
//
// tc_conv.cuh fetches the voxel brick once per tap: every input voxel crosses L2->SM 27 times and the kernel is bound by
// L2 request bandwidth (measured 16-35 B/clk/SM).  Here the output tile is ONE d-plane of 16(h) x 8(w) voxels and, per
// kd, one TMA box brings the (16+2) x (8+2) halo plane [18][10][C] into swizzled smem.  All nine (kh,kw) taps of that kd
// are then just *descriptor views* of the same tile:
//       start = tile + (kh*10 + kw) * row_bytes ,   SBO (next 8-row group = next h) = 10 * row_bytes
// which is legal because UMMA (like TMA) applies the 32/64/128-byte swizzle XOR to absolute smem address bits (the same
// property the +32 B K-advance inside a swizzle atom relies on).  3 TMA loads and zero copies per tile instead of 27.
#pragma once
#include "tc_conv.cuh"

namespace b200 {
namespace tc {

static constexpr int HTH = 16, HTW = 8, HALO_H = HTH + 2, HALO_W = HTW + 2;

struct HaloParams {
  int N, D, H, W, Ci, Co;
  int kc, row_bytes, nchunk;     // channels per chunk (<= 64), smem row size, Ci/kc
  int tiles_w, tiles_h; long total_tiles;
  int halo_bytes;                // HALO_H*HALO_W*row_bytes rounded up to 1024
  int b_bytes;                   // one (tap, chunk) weight tile, rounded to 1024
  int resident;                  // all 27*nchunk weight tiles stay in smem
  int stage_bytes, stages; uint32_t tmem_cols;
  bf16* out; int pitch, coff, accumulate; double* stats;
};

__global__ void __launch_bounds__(192, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t wres_bytes = p.resident ? (uint32_t)(27 * p.nchunk) * p.b_bytes : 0;
  uint8_t* ring = smem + wres_bytes;
  uint64_t* full = (uint64_t*)(ring + (size_t)p.stages * p.stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* wfull = tempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(wfull + 1);
  float* red = (float*)(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nstage_per_tile = 3 * p.nchunk;   // (kd, chunk)
  const uint32_t w_tx = (uint32_t)(p.Co * p.row_bytes);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
    mbar_init(wfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.resident) {   // weight tile index = (tap * nchunk + chunk), tap = (kd*3+kh)*3+kw
      mbar_expect_tx(wfull, (uint32_t)(27 * p.nchunk) * w_tx);
      for (int tap = 0; tap < 27; ++tap)
        for (int ch = 0; ch < p.nchunk; ++ch)
          tma_load_2d(smem_u32(smem) + (uint32_t)(tap * p.nchunk + ch) * p.b_bytes, &map_w, wfull, ch * p.kc, tap * p.Co);
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

  if (warp == 0) {
    // ---- TMA producer: one halo plane (+ its 9 weight tiles when not resident) per (kd, chunk)
    int stage = 0; uint32_t phase = 0;
    const uint32_t ring_u = smem_u32(ring);
    const uint32_t tx = (uint32_t)(HALO_H * HALO_W * p.row_bytes) + (p.resident ? 0u : 9u * w_tx);
    for (long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int tw = (int)(t % p.tiles_w); long r = t / p.tiles_w; int th = (int)(r % p.tiles_h); r /= p.tiles_h;
      int d = (int)(r % p.D); int n = (int)(r / p.D);
      for (int kd = 0; kd < 3; ++kd)
        for (int ch = 0; ch < p.nchunk; ++ch) {
          mbar_wait(empty + stage, phase ^ 1);
          if (elect_one()) {
            const uint32_t base = ring_u + (uint32_t)stage * p.stage_bytes;
            mbar_expect_tx(full + stage, tx);
            tma_load_5d(base, &map_x, full + stage, ch * p.kc, tw * HTW - 1, th * HTH - 1, d + kd - 1, n);
            if (!p.resident)
              for (int j = 0; j < 9; ++j)
                tma_load_2d(base + p.halo_bytes + j * p.b_bytes, &map_w, full + stage, ch * p.kc, (kd * 9 + j) * p.Co);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Co >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t layout = p.row_bytes == 128 ? 2u : (p.row_bytes == 64 ? 4u : 6u);
    const uint32_t a_hi = desc_hi((uint32_t)(HALO_W * p.row_bytes), layout);   // next 8-row group = next h line of the halo
    const uint32_t b_hi = desc_hi(8 * p.row_bytes, layout);
    const uint32_t ring_u = smem_u32(ring);
    const uint32_t a_lo0 = desc_lo(ring_u, 16);
    const uint32_t b_lo0 = p.resident ? desc_lo(smem_u32(smem), 16) : desc_lo(ring_u + p.halo_bytes, 16);
    const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4, b_units = (uint32_t)p.b_bytes >> 4, row_units = (uint32_t)p.row_bytes >> 4;
    const int ksteps = p.kc / 16;
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    for (long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait(tempty + acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.Co);
      int s = 0;
      for (int kd = 0; kd < 3; ++kd)
        for (int ch = 0; ch < p.nchunk; ++ch, ++s) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t a_st = a_lo0 + (uint32_t)stage * stage_units;
          const uint32_t b_st = p.resident ? b_lo0 + (uint32_t)((kd * 9) * p.nchunk + ch) * b_units : b_lo0 + (uint32_t)stage * stage_units;
          const uint32_t b_tap = p.resident ? (uint32_t)p.nchunk * b_units : b_units;   // distance between consecutive taps' weight tiles
          if (elect_one()) {
            uint32_t b_lo = b_st;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const uint32_t a_lo = a_st + (uint32_t)(kh * HALO_W + kw) * row_units;
                for (int k = 0; k < ksteps; ++k)
                  umma_f16(tmem_d, desc64(a_lo + 2 * k, a_hi), desc64(b_lo + 2 * k, b_hi), idesc, (s | kh | kw | k) ? 1u : 0u);
                b_lo += b_tap;
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
    const int q = warp & 3, ew = warp - 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int tw = (int)(t % p.tiles_w); long r = t / p.tiles_w; int th = (int)(r % p.tiles_h); r /= p.tiles_h;
      int d = (int)(r % p.D); int n = (int)(r / p.D);
      const int row = q * 32 + lane;
      const int w = tw * HTW + (row & 7), h = th * HTH + (row >> 3);
      const bool valid = (w < p.W) && (h < p.H);
      bf16* dst = p.out + ((((long)n * p.D + d) * p.H + h) * p.W + w) * p.pitch + p.coff;
      mbar_wait(tfull + acc, acc_phase);
      tc_fence_after();
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
          Vec16<bf16> o0, o1;
#pragma unroll
          for (int j = 0; j < 8; ++j) { o0.v[j] = v[j]; o1.v[j] = v[8 + j]; }
          o0.store(dst + c0); o1.store(dst + c0 + 8);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
      if (p.stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        int e = (warp - 2) * 32 + lane;
        for (int i = e; i < 2 * p.Co; i += 128) {
          float tot = red[i] + red[2 * p.Co + i] + red[4 * p.Co + i] + red[6 * p.Co + i];
          int c = i % p.Co, which = i / p.Co;
          atomicAdd(p.stats + ((long)n * p.Co + c) * 2 + which, (double)tot);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// does the (kd, chunk) stage (halo plane + 9 weight tiles unless all weights are resident) fit at least twice?
static inline bool conv_halo_supported(int Ci, int Co) {
  int kc = Ci % 64 == 0 ? 64 : (Ci % 32 == 0 ? 32 : 16), rb = kc * 2, nchunk = Ci / kc;
  int halo = ((HALO_H * HALO_W * rb + 1023) / 1024) * 1024, b = ((Co * rb + 1023) / 1024) * 1024;
  bool resident = (long)27 * nchunk * b <= 112 * 1024;
  int stage = halo + (resident ? 0 : 9 * b), budget = 200 * 1024 - (resident ? 27 * nchunk * b : 0);
  return Ci % 16 == 0 && Co % 16 == 0 && Co <= 256 && budget / stage >= 2;
}

// 3x3x3 only.  wp: packed bf16 [27][Co][Ci] (same packing as tc::conv).
static int conv_halo(const bf16* x, int in_pitch, int in_coff, int Ci, int N, int D, int H, int W, const bf16* wp, int Co,
                     bf16* out, int out_pitch, int out_coff, int accumulate, double* stats, cudaStream_t st) {
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
  HaloParams p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co;
  p.kc = Ci % 64 == 0 ? 64 : (Ci % 32 == 0 ? 32 : 16);
  p.row_bytes = p.kc * 2; p.nchunk = Ci / p.kc;
  p.tiles_w = cdiv(W, HTW); p.tiles_h = cdiv(H, HTH);
  p.total_tiles = (long)N * D * p.tiles_h * p.tiles_w;
  p.halo_bytes = ((HALO_H * HALO_W * p.row_bytes + 1023) / 1024) * 1024;
  p.b_bytes = ((Co * p.row_bytes + 1023) / 1024) * 1024;
  p.resident = ((long)27 * p.nchunk * p.b_bytes <= 112 * 1024) ? 1 : 0;
  p.stage_bytes = p.halo_bytes + (p.resident ? 0 : 9 * p.b_bytes);
  int budget = 200 * 1024 - (p.resident ? 27 * p.nchunk * p.b_bytes : 0);
  p.stages = budget / p.stage_bytes; if (p.stages > 8) p.stages = 8;
  B200_CHECK(p.stages >= 2, "halo conv smem budget exceeded (Ci=%d Co=%d)", Ci, Co);
  uint32_t cols = 2 * Co, pw = 32; while (pw < cols) pw <<= 1; p.tmem_cols = pw;
  p.out = out; p.pitch = out_pitch; p.coff = out_coff; p.accumulate = accumulate; p.stats = stats;

  CUtensorMapSwizzle sw = p.row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap mx, mw;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Ci, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)W * in_pitch * 2, (cuuint64_t)H * W * in_pitch * 2, (cuuint64_t)D * H * W * in_pitch * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.kc, HALO_W, HALO_H, 1, 1};
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
  size_t smem = (size_t)p.stages * p.stage_bytes + (p.resident ? (size_t)27 * p.nchunk * p.b_bytes : 0) + 1024 + 256 + 8 * Co * sizeof(float) + 64;
  B200_CHECK(smem <= 227 * 1024, "halo conv smem budget exceeded (%zu)", smem);
  static bool attr_done = false;
  if (!attr_done) { B200_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_done = true; }
  int grid = (int)(p.total_tiles < num_sms() ? p.total_tiles : num_sms());
  conv_halo_kernel<<<grid, 192, smem, st>>>(mx, mw, p);
  B200_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc
}  // namespace b200
