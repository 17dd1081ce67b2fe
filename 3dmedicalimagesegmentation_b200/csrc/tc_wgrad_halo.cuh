// tcgen05 weight gradient of the 3x3x3 convolution for 16/32-channel inputs, with on-SM halo reuse:
//
//   dW[co][ci][(kd,kh,kw)] = sum_v dy[v, co] * x[v + (kd,kh,kw) - 1, ci]
//
// tc_wgrad.cuh fetches the x brick once per tap (27 TMA boxes of 128 32-byte rows per 128 voxels: TMA-row bound).  Here a
// pipeline stage holds the three (16+2) x (8+2) halo planes of x (ONE depth-3 TMA box, as in tc_conv_halo.cuh) plus the
// 16 x 8 dy tile.  Voxels are the reduction dimension, both operands are MN-major smem views:
//   A (M side): x.  The M-atoms of one MMA are the kw taps: atom j starts one voxel (= one smem row) after atom j-1, so
//               LBO = row_bytes; 8 k-rows = 8 consecutive w of one h line (contiguous), SBO = 10 rows (next h line).
//               M = 128 = (128 / Ci) atoms of Ci channels; atoms kw >= 3 are junk rows the epilogue never reads.
//   B (N side): the dy tile, N = Co.
//   D         : nine fp32 accumulators [128 x Co] in TMEM, one per (kd,kh), resident across ALL tiles of the CTA; one
//               atomic epilogue at the end.
// Per 128 voxels: 1 + 1 TMA boxes and 72 MMAs (9 (kd,kh) x 8 k-steps of 16 voxels), issued from a fully unrolled loop.
//
// kh stacked along N (`stack`, default): these N = 16 MMAs are bound by their fixed issue cost (ncu: tensor pipe 96 % busy at 72 MMAs
// of 64 x 16 x 16 per tile), so the kh shift is moved from x to dy --
//       dW[kd,kh,kw] = sum_b x[b + (kd-1, 0, kw-1)] * dy[b - (0, kh-1, 0)]        (b = v + the kh shift; out-of-volume terms are zero)
// -- and becomes the N-atom index of the B operand exactly as kw is the M-atom index of A: the dy tile is fetched with one halo line
// above and below (18 x 8 voxels), N-atom i starts i lines further down (LBO = SBO = one line) and holds kh = 2 - i.  x then needs its
// halo in w and d only (16 x 10 voxels per plane).  One MMA is M x 3*Co x 16 and a tile takes 24 of them instead of 72; the three
// accumulators (one per kd) are [M x 3*Co].
#pragma once
#include "tc_conv_halo.cuh"
#include "tc_wgrad.cuh"

namespace b200 {
namespace tc {

struct WgradHaloParams {
  int N, D, H, W, Ci, Co;
  int tiles_w, tiles_h, tiles_per_n, total_tiles;
  int plane_bytes, x_bytes, dy_bytes, stage_bytes, stages; uint32_t tmem_cols;
  float* dW;
  int stack;                // kh stacked along N (see the header)
  float* dW3;               // optional: weight gradient [Co][Ci] of the residual block's 1x1x1 convolution on the same x (fused: a tenth
                            // accumulator fed by the centre tap and a second dy tile); nullptr = plain 3x3x3 weight gradient
  long long* trace;
};

template <int CI>   // 16 or 32
__global__ void __launch_bounds__(192, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_dy3,
                  const WgradHaloParams p) {
  constexpr uint32_t ROW_BYTES = CI * 2, ROW_UNITS = ROW_BYTES / 16;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* done = empty + p.stages;
  uint32_t* tmem_slot = (uint32_t*)(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  trace_start(p.trace);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ---- TMA producer
    int stage = 0; uint32_t phase = 0;
    const uint32_t smem_u = smem_u32(smem);
    const uint32_t tx = (uint32_t)(3 * p.plane_bytes) + (uint32_t)((p.stack ? 144 : 128) * p.Co * 2) + (p.dW3 ? (uint32_t)(128 * p.Co * 2) : 0u);
    int t = blockIdx.x;
    int n = t / p.tiles_per_n, r = t - n * p.tiles_per_n;
    for (; t < p.total_tiles; t += gridDim.x) {
      const int tw = r % p.tiles_w, q = r / p.tiles_w, th = q % p.tiles_h, d = q / p.tiles_h;
      mbar_wait(empty + stage, phase ^ 1);
      if (elect_one()) {
        const uint32_t base = smem_u + (uint32_t)stage * p.stage_bytes;
        mbar_expect_tx(full + stage, tx);
        tma_load_5d(base, &map_x, full + stage, 0, tw * HTW - 1, th * HTH - (p.stack ? 0 : 1), d - 1, n);
        tma_load_5d(base + p.x_bytes, &map_dy, full + stage, 0, tw * HTW, th * HTH - (p.stack ? 1 : 0), d, n);
        if (p.dW3) tma_load_5d(base + p.x_bytes + p.dy_bytes, &map_dy3, full + stage, 0, tw * HTW, th * HTH, d, n);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
      r += gridDim.x;
      while (r >= p.tiles_per_n) { r -= p.tiles_per_n; ++n; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: both operands MN-major (bits 15, 16)
    // CI = 16: M = 64 (4 kw atoms, 3 useful) halves the A-operand smem read that bounds these N = 16 MMAs (ncu: tensor pipe 67 %
    // busy at M = 128 with 5 junk atoms of 8); CI = 32: M = 128 = 4 atoms.
    constexpr uint32_t MROWS = CI == 16 ? 64u : 128u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Co >> 3) << 17) | ((MROWS >> 4) << 24);
    const uint32_t a_layout = CI == 32 ? 4u : 6u;                              // 64 B / 32 B swizzle
    const uint32_t b_row_bytes = (uint32_t)p.Co * 2, b_layout = b_row_bytes == 64 ? 4u : 6u;
    const uint32_t a_hi = desc_hi(HALO_W * ROW_BYTES, a_layout);               // SBO: next 8-k-row group = next h line
    const uint32_t b_hi = desc_hi(8 * b_row_bytes, b_layout);
    const uint32_t smem_u = smem_u32(smem);
    const uint32_t a_lo0 = desc_lo(smem_u, ROW_BYTES);                         // LBO: next kw atom = next voxel row
    const uint32_t b_lo0 = desc_lo(smem_u + (uint32_t)p.x_bytes, 16);
    const uint32_t stage_units = (uint32_t)p.stage_bytes >> 4, plane_units = (uint32_t)p.plane_bytes >> 4;
    const uint32_t b_kstep = (16u * b_row_bytes) >> 4;                         // 16 voxels = 2 h lines of the dy tile
    constexpr uint32_t A_KSTEP = 2 * HALO_W * ROW_UNITS;                       // ... = 2 h lines of the halo
    const uint32_t co = (uint32_t)p.Co;
    int stage = 0; uint32_t phase = 0; uint32_t accum = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      mbar_wait(full + stage, phase);
      tc_fence_after();
      const uint32_t a_st = a_lo0 + (uint32_t)stage * stage_units, b_st = b_lo0 + (uint32_t)stage * stage_units;
      if (p.stack) {
        if (elect_one()) {
          // B: three kh atoms of Co channels, one dy line apart (LBO = SBO = 8 rows)
          const uint32_t idesc3 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((3 * p.Co) >> 3) << 17) | ((MROWS >> 4) << 24);
          const uint32_t b_st3 = desc_lo(smem_u + (uint32_t)p.x_bytes + (uint32_t)stage * (uint32_t)p.stage_bytes, 8 * b_row_bytes);
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            const uint32_t a_t = a_st + (uint32_t)kd * plane_units;
            const uint32_t tmem_d = tmem_base + (uint32_t)kd * 3u * co;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              umma_f16(tmem_d, desc64(a_t + (uint32_t)j * A_KSTEP, a_hi), desc64(b_st3 + (uint32_t)j * b_kstep, b_hi), idesc3, j == 0 ? accum : 1u);
          }
          if (p.dW3) {   // 1x1x1: centre plane of x against the second dy tile (no halo); the kw = 1 atom is the result
            const uint32_t a_t = a_st + plane_units;
            const uint32_t b3 = b_st + ((uint32_t)p.dy_bytes >> 4);
            const uint32_t tmem_d = tmem_base + 9u * co;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              umma_f16(tmem_d, desc64(a_t + (uint32_t)j * A_KSTEP, a_hi), desc64(b3 + (uint32_t)j * b_kstep, b_hi), idesc, j == 0 ? accum : 1u);
          }
          umma_commit(empty + stage);
        }
      } else if (elect_one()) {
#pragma unroll
        for (int kd = 0; kd < 3; ++kd)
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t a_t = a_st + (uint32_t)kd * plane_units + (uint32_t)(kh * HALO_W) * ROW_UNITS;
            const uint32_t tmem_d = tmem_base + (uint32_t)(kd * 3 + kh) * co;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              umma_f16(tmem_d, desc64(a_t + (uint32_t)j * A_KSTEP, a_hi), desc64(b_st + (uint32_t)j * b_kstep, b_hi), idesc, j == 0 ? accum : 1u);
          }
        if (p.dW3) {   // 1x1x1: centre (kd, kh) = (1, 1) view of x against the second dy tile; the kw = 1 atom is the result
          const uint32_t a_t = a_st + plane_units + (uint32_t)HALO_W * ROW_UNITS;
          const uint32_t b3 = b_st + ((uint32_t)p.dy_bytes >> 4);
          const uint32_t tmem_d = tmem_base + 9u * co;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_f16(tmem_d, desc64(a_t + (uint32_t)j * A_KSTEP, a_hi), desc64(b3 + (uint32_t)j * b_kstep, b_hi), idesc, j == 0 ? accum : 1u);
        }
        umma_commit(empty + stage);
      }
      __syncwarp();
      accum = 1u;
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  } else {
    // ---- epilogue: lane row -> (kw atom, ci); atoms kw >= 3 are junk
    const int q = warp & 3;
    if ((int)blockIdx.x < p.total_tiles) {
      mbar_wait(done, 0);
      tc_fence_after();
      // M = 128: accumulator row = TMEM lane.  M = 64: rows 16q..16q+15 sit in lanes 32q..32q+15 (16 lanes per warp quarter).
      const int row = CI == 16 ? (lane < 16 ? q * 16 + lane : 1 << 20) : q * 32 + lane;
      const int kw = row / CI, ci = row % CI;
      for (int a = 0; a < 9; ++a) {
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * p.Co);
        for (int c0 = 0; c0 < p.Co; c0 += 16) {
          float v[16];
          tmem_ld16(trow + c0, v);
          if (kw < 3) {
            const int tap3 = p.stack ? (a / 3) * 3 + (2 - a % 3) : a;      // (kd, kh) of this accumulator block
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(p.dW + ((long)(c0 + j) * p.Ci + ci) * 27 + tap3 * 3 + kw, v[j]);
          }
        }
      }
      if (p.dW3) {
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(9 * p.Co);
        for (int c0 = 0; c0 < p.Co; c0 += 16) {
          float v[16];
          tmem_ld16(trow + c0, v);
          if (kw == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(p.dW3 + (long)(c0 + j) * p.Ci + ci, v[j]);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

static inline bool wgrad_halo_supported(int Ci, int Co, int ks) {
  return ks == 3 && (Ci == 16 || Ci == 32) && (Co == 16 || Co == 32) && !getenv("B200_NO_WGRAD_HALO");
}

template <int CI>
static int wgrad_halo_launch(const CUtensorMap& mx, const CUtensorMap& mdy, const CUtensorMap& mdy3, const WgradHaloParams& p, int grid, size_t smem,
                             cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) { B200_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel<CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_done = true; }
  B200_CUDA(launch_pdl(wgrad_halo_kernel<CI>, dim3(grid), dim3(192), smem, st, mx, mdy, mdy3, p));
  B200_LAUNCH_CHECK();
  return 0;
}

// dW (fp32 [Co][Ci][27]) must be zero on entry.  dy3 / dW3 (optional, dW3 fp32 [Co][Ci] zero on entry): the 1x1x1 convolution of the
// residual block reads the same x, so its weight gradient rides along as a tenth accumulator (one more TMA box, 8 more MMAs per tile).
static int conv_wgrad_halo(const bf16* x, int x_pitch, int x_coff, int Ci, const bf16* dy, int dy_pitch, int dy_coff, int Co, int N, int D, int H, int W,
                           float* dW, cudaStream_t st, const bf16* dy3 = nullptr, int dy3_pitch = 0, int dy3_coff = 0, float* dW3 = nullptr) {
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
  WgradHaloParams p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co;
  p.tiles_w = cdiv(W, HTW); p.tiles_h = cdiv(H, HTH); p.tiles_per_n = D * p.tiles_h * p.tiles_w;
  long total = (long)N * p.tiles_per_n;
  B200_CHECK(total < (1L << 30), "wgrad halo: too many tiles");
  p.total_tiles = (int)total;
  const int rb = Ci * 2;
  static const bool old_form = getenv("B200_WGRAD_HALO_OLD") != nullptr;
  p.stack = old_form ? 0 : 1;
  p.plane_bytes = (p.stack ? HTH : HALO_H) * HALO_W * rb;
  // the junk atoms of the last k-step read up to (128/Ci - 3) voxel rows past the third plane: keep them inside the stage
  p.x_bytes = ((3 * p.plane_bytes + (128 / Ci) * rb + 1023) / 1024) * 1024;
  p.dy_bytes = (((p.stack ? 144 : 128) * Co * 2 + 1023) / 1024) * 1024;      // stack: 18 lines of 8 voxels (one halo line above and below)
  p.dW3 = (dy3 && dW3) ? dW3 : nullptr;
  p.stage_bytes = p.x_bytes + p.dy_bytes + (p.dW3 ? ((128 * Co * 2 + 1023) / 1024) * 1024 : 0);
  p.stages = (200 * 1024) / p.stage_bytes; if (p.stages > 8) p.stages = 8;
  B200_CHECK(p.stages >= 2, "wgrad halo smem budget exceeded");
  uint32_t cols = (p.dW3 ? 10 : 9) * Co, pw = 32; while (pw < cols) pw <<= 1; p.tmem_cols = pw;
  B200_CHECK(p.tmem_cols <= 512, "wgrad halo TMEM budget exceeded");
  p.dW = dW;
  p.trace = trace_slot(); if (p.trace) trace_tag("wgrad_halo %dx%d @%d%s", Ci, Co, D, p.dW3 ? " +1x1" : "");
  CUtensorMap mx, mdy;
  {
    CUtensorMapSwizzle sw = rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint64_t dims[5] = {(cuuint64_t)Ci, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)x_pitch * 2, (cuuint64_t)W * x_pitch * 2, (cuuint64_t)H * W * x_pitch * 2, (cuuint64_t)D * H * W * x_pitch * 2};
    cuuint32_t box[5] = {(cuuint32_t)Ci, HALO_W, (cuuint32_t)(p.stack ? HTH : HALO_H), 3, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(x + x_coff), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "wgrad halo x tensor map failed (%d)", (int)r);
  }
  {
    CUtensorMapSwizzle sw = Co * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint64_t dims[5] = {(cuuint64_t)Co, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)dy_pitch * 2, (cuuint64_t)W * dy_pitch * 2, (cuuint64_t)H * W * dy_pitch * 2, (cuuint64_t)D * H * W * dy_pitch * 2};
    cuuint32_t box[5] = {(cuuint32_t)Co, HTW, (cuuint32_t)(p.stack ? HALO_H : HTH), 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&mdy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(dy + dy_coff), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "wgrad halo dy tensor map failed (%d)", (int)r);
  }
  CUtensorMap mdy3 = mdy;
  if (p.dW3) {
    CUtensorMapSwizzle sw = Co * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint64_t dims[5] = {(cuuint64_t)Co, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)dy3_pitch * 2, (cuuint64_t)W * dy3_pitch * 2, (cuuint64_t)H * W * dy3_pitch * 2, (cuuint64_t)D * H * W * dy3_pitch * 2};
    cuuint32_t box[5] = {(cuuint32_t)Co, HTW, HTH, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&mdy3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(dy3 + dy3_coff), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_CHECK(r == CUDA_SUCCESS, "wgrad halo dy3 tensor map failed (%d)", (int)r);
  }
  size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + 256;
  int grid = (int)(p.total_tiles < num_sms() ? p.total_tiles : num_sms());
  if (Ci == 16) return wgrad_halo_launch<16>(mx, mdy, mdy3, p, grid, smem, st);
  return wgrad_halo_launch<32>(mx, mdy, mdy3, p, grid, smem, st);
}

}  // namespace tc
}  // namespace b200
