// tcgen05 weight gradient of the 3-D convolution (k in {1,3}, stride 1, same padding), channels-last bf16:
//
//   dW[co][ci][tap] = sum_v dy[v, co] * x[v + tap - pad, ci]
//
// GEMM view with the VOXELS as the reduction dimension.  Both operands are MN-major straight out of TMA: the smem
// image of a 4x4x8 voxel brick is [128 voxel rows][channels], i.e. K rows of MN-contiguous elements.
//   A (M side) : shifted x bricks.  An "atom" is (tap, <=64-channel chunk); 128/atom_ch atoms are stacked along M
//                (LBO = one brick) so that one MMA is 128 x Co x 16 although a layer has only 16..64 channels.
//   B (N side) : the dy brick, N = Co.
//   D          : fp32 in TMEM, one [128 x Co] accumulator per M-tile, resident across ALL voxel bricks this CTA owns;
//                a single epilogue at the end hands the CTA's partial dW out.
// Work split: grid = (M-tile groups that fit 512 TMEM columns) x (voxel-brick splits).
// Epilogue: with a scratch buffer the partial tiles [split][M rows][Co] are written as swizzled 32 x 32 boxes through TMA stores and
// wgrad_reduce_kernel sums the splits in a fixed order into the PyTorch layout dW[co][ci][tap] -- deterministic, and the 24^3 / 12^3
// layers no longer pay one fp32 atomic per (split, element): 8 M atomics for a 0.9 M-element dW at 256 x 128 @12^3 (measured 66 us for
// 6 GFLOP, the MMAs being ~5 us of it).  Without scratch: fp32 atomics as before.
#pragma once
#include "tc_conv.cuh"

namespace b200 {
namespace tc {

struct WgradParams {
  int N, D, H, W, Ci, Co, ks, taps;
  int atom_ch, a_row_bytes, a_atom_bytes;       // A atom: [128 voxels][atom_ch]
  int atoms_per_tap, atoms_per_tile, total_atoms, n_mtiles, mt_per_cta, n_groups, vsplit;
  int b_ch, b_row_bytes, b_chunks, b_chunk_bytes; // dy brick in <=64-channel chunks
  int tiles_w, tiles_h, tiles_d; long total_tiles;
  int stages; uint32_t tmem_cols;
  float* dW;
  float* scratch;           // [vsplit][n_mtiles * 128][Co] partial tiles (nullptr: atomic epilogue)
  long long* trace;
};

__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr, uint32_t lbo_bytes, int row_bytes) {
  // MN-major canonical layout: atoms of `row_bytes` along MN (LBO apart), 8-k-row groups 8*row_bytes apart (SBO)
  uint64_t lt = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(((uint32_t)(8 * row_bytes) >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= lt << 61;
  return d;
}

static __global__ void __launch_bounds__(192, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_s,
             const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_stage_bytes = (uint32_t)p.atoms_per_tile * p.a_atom_bytes;   // 32 KB
  const uint32_t b_bytes = (uint32_t)p.b_chunks * p.b_chunk_bytes;
  uint8_t* smem_b = smem + (size_t)p.stages * a_stage_bytes;
  uint64_t* afull = (uint64_t*)(smem_b + 2 * (size_t)b_bytes);
  uint64_t* aempty = afull + p.stages;
  uint64_t* bfull = aempty + p.stages;   // [2]
  uint64_t* bempty = bfull + 2;          // [2]
  uint64_t* done = bempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.x % p.n_groups, vs = blockIdx.x / p.n_groups;
  const int mt0 = grp * p.mt_per_cta, mt1 = min(p.n_mtiles, mt0 + p.mt_per_cta);
  const int pad = p.ks >> 1;
  trace_start(p.trace);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(afull + s, 1); mbar_init(aempty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bfull + s, 1); mbar_init(bempty + s, 1); }
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
    // ---- TMA producer: warp-uniform loop, one elected lane issues
    int stage = 0; uint32_t phase = 0; int bb = 0; uint32_t bphase = 0;
    const uint32_t smem_a_u = smem_u32(smem), smem_b_u = smem_u32(smem_b);
    for (long t = vs; t < p.total_tiles; t += p.vsplit) {
      int tw = (int)(t % p.tiles_w); long r = t / p.tiles_w; int th = (int)(r % p.tiles_h); r /= p.tiles_h;
      int td = (int)(r % p.tiles_d); int n = (int)(r / p.tiles_d);
      const int w0 = tw * TW, h0 = th * TH, d0 = td * TD;
      mbar_wait(bempty + bb, bphase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(bfull + bb, b_bytes);
        for (int c = 0; c < p.b_chunks; ++c)
          tma_load_5d(smem_b_u + (uint32_t)bb * b_bytes + (uint32_t)c * p.b_chunk_bytes, &map_dy, bfull + bb, c * p.b_ch, w0, h0, d0, n);
      }
      __syncwarp();
      if (++bb == 2) { bb = 0; bphase ^= 1; }
      // running atom state: atom a = tap*atoms_per_tap + ch, starting at the first atom of this CTA's M-tiles
      int a = mt0 * p.atoms_per_tile;
      int tap = a / p.atoms_per_tap, ch = a - tap * p.atoms_per_tap;
      int kw = tap % p.ks, kh = (tap / p.ks) % p.ks, kd = tap / (p.ks * p.ks);
      for (int mt = mt0; mt < mt1; ++mt) {
        const int cnt = min(p.atoms_per_tile, p.total_atoms - mt * p.atoms_per_tile);
        mbar_wait(aempty + stage, phase ^ 1);
        const uint32_t base = smem_a_u + (uint32_t)stage * a_stage_bytes;
        const bool leader = elect_one();
        if (leader) mbar_expect_tx(afull + stage, (uint32_t)cnt * p.a_atom_bytes);
        for (int i = 0; i < cnt; ++i) {
          if (leader) tma_load_5d(base + i * p.a_atom_bytes, &map_x, afull + stage, ch * p.atom_ch, w0 + kw - pad, h0 + kh - pad, d0 + kd - pad, n);
          if (++ch == p.atoms_per_tap) { ch = 0; if (++kw == p.ks) { kw = 0; if (++kh == p.ks) { kh = 0; ++kd; } } }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer; both operands MN-major (bits 15, 16)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Co >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto layout_of = [](int rb) { return rb == 128 ? 2u : (rb == 64 ? 4u : 6u); };
    const uint32_t a_hi = desc_hi(8 * p.a_row_bytes, layout_of(p.a_row_bytes)), b_hi = desc_hi(8 * p.b_row_bytes, layout_of(p.b_row_bytes));
    const uint32_t a_lo0 = desc_lo(smem_u32(smem), p.a_atom_bytes), b_lo0 = desc_lo(smem_u32(smem_b), p.b_chunk_bytes);
    const uint32_t a_kstep = (16u * p.a_row_bytes) >> 4, b_kstep = (16u * p.b_row_bytes) >> 4;   // 16 voxels per MMA
    const uint32_t a_stage_units = a_stage_bytes >> 4, b_units = b_bytes >> 4;
    int stage = 0; uint32_t phase = 0; int bb = 0; uint32_t bphase = 0; uint32_t accum = 0;
    for (long t = vs; t < p.total_tiles; t += p.vsplit) {
      mbar_wait(bfull + bb, bphase);
      tc_fence_after();
      const uint32_t b_lo = b_lo0 + (uint32_t)bb * b_units;
      for (int mt = mt0; mt < mt1; ++mt) {
        mbar_wait(afull + stage, phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)stage * a_stage_units;
        const uint32_t tmem_d = tmem_base + (uint32_t)((mt - mt0) * p.Co);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 8; ++j)   // 8 x 16 voxels = one brick
            umma_f16(tmem_d, desc64(a_lo + j * a_kstep, a_hi), desc64(b_lo + j * b_kstep, b_hi), idesc, (j == 0) ? accum : 1u);
          umma_commit(aempty + stage);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(bempty + bb);
      __syncwarp();
      if (++bb == 2) { bb = 0; bphase ^= 1; }
      accum = 1u;
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const bool any = vs < p.total_tiles;
    if (any) {
      mbar_wait(done, 0);
      tc_fence_after();
      const int row = q * 32 + lane;
      int sbuf = 0;
      for (int mt = mt0; mt < mt1; ++mt) {
        int a = mt * p.atoms_per_tile + row / p.atom_ch;
        int tap = a / p.atoms_per_tap, ci = (a - tap * p.atoms_per_tap) * p.atom_ch + row % p.atom_ch;
        const bool valid = a < p.total_atoms;
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((mt - mt0) * p.Co);
        if (p.scratch) {
          // rows [mt*128 + 32q, +32) x 32 columns per box; the A-stage ring is free (all MMAs retired): boxes live in its first 32 KB
          uint8_t* wstg = smem + (uint32_t)q * 2u * 4096u;
          for (int c0 = 0; c0 < p.Co; c0 += 32) {
            float v[32];
            tmem_ld16x2(trow + c0, v, true);
            uint8_t* box = wstg + (uint32_t)sbuf * 4096u;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
            stage_row32(box, lane, v, 0.f);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { tma_store_3d(&map_s, smem_u32(box), c0, mt * 128 + q * 32, vs); bulk_commit(); }
            sbuf ^= 1;
          }
          continue;
        }
        for (int c0 = 0; c0 < p.Co; c0 += 16) {
          float v[16];
          tmem_ld16(trow + c0, v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(p.dW + ((long)(c0 + j) * p.Ci + ci) * p.taps + tap, v[j]);
          }
        }
      }
      if (p.scratch && lane == 0) bulk_wait_read<0>();
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

// dW[co][ci][tap] = sum over splits, in split order, of scratch[split][row][co]; row -> (tap, ci) as in the kernel's M-tile layout
static __global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ scratch, int vsplit, int rows, int Co, int Ci, int taps, int atom_ch,
                                                                  int atoms_per_tap, int total_atoms, float* __restrict__ dW) {
  pdl_wait();
  const long total = (long)rows * Co;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int co = (int)(e % Co); const int r = (int)(e / Co);
    const int a = r / atom_ch;
    if (a >= total_atoms) continue;
    const int tap = a / atoms_per_tap, ci = (a - tap * atoms_per_tap) * atom_ch + r % atom_ch;
    float acc = 0.f;
    for (int sp = 0; sp < vsplit; ++sp) acc += scratch[(long)sp * total + e];
    dW[((long)co * Ci + ci) * taps + tap] = acc;
  }
}
// bytes of scratch the deterministic epilogue needs for this layer (0: layer not taken by this kernel)
static inline size_t wgrad_scratch_bytes(int Ci, int Co, int ks, int N, int D, int H, int W);

static inline bool wgrad_supported(int Ci, int Co, int x_pitch, int x_coff, int dy_pitch, int dy_coff) {
  auto ok = [](int c) { return c == 16 || c == 32 || (c % 64 == 0 && c <= 256); };
  return ok(Ci) && ok(Co) && Co <= 128 && x_pitch % 8 == 0 && x_coff % 8 == 0 && dy_pitch % 8 == 0 && dy_coff % 8 == 0;
}

static int make_brick_map(CUtensorMap* m, const bf16* base, int pitch, int C, int box_c, int N, int D, int H, int W) {
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
  int row_bytes = box_c * 2;
  CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2, (cuuint64_t)D * H * W * pitch * 2};
  cuuint32_t box[5] = {(cuuint32_t)box_c, TW, TH, TD, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, "brick tensor map failed (%d): C=%d pitch=%d dims %dx%dx%d", (int)r, C, pitch, D, H, W);
  return 0;
}

// dW (fp32 [Co][Ci][taps]) must be zero on entry.
static void wgrad_plan(WgradParams& p, int Ci, int Co, int ks, int N, int D, int H, int W);
static inline size_t wgrad_scratch_bytes(int Ci, int Co, int ks, int N, int D, int H, int W) {
  WgradParams p; wgrad_plan(p, Ci, Co, ks, N, D, H, W);
  // 1x1x1 layers keep the atomic epilogue: their dW is a single M-tile and the second launch costs more than the atomics (measured)
  return (Co % 32 || ks != 3) ? 0 : (size_t)p.vsplit * p.n_mtiles * 128 * Co * sizeof(float);
}
static int conv_wgrad(const bf16* x, int x_pitch, int x_coff, int Ci, const bf16* dy, int dy_pitch, int dy_coff, int Co, int N, int D, int H, int W,
                      int ks, float* dW, cudaStream_t st, float* scratch = nullptr, size_t scratch_bytes = 0) {
  WgradParams p;
  wgrad_plan(p, Ci, Co, ks, N, D, H, W);
  static const bool atomic_only = getenv("B200_WGRAD_ATOMIC") != nullptr;
  const size_t need = wgrad_scratch_bytes(Ci, Co, ks, N, D, H, W);
  p.scratch = (!atomic_only && scratch && need && need <= scratch_bytes) ? scratch : nullptr;
  uint32_t a_stage = (uint32_t)p.atoms_per_tile * p.a_atom_bytes, b_bytes = (uint32_t)p.b_chunks * p.b_chunk_bytes;
  B200_CHECK(p.tmem_cols <= 512, "wgrad TMEM budget exceeded");
  B200_CHECK(p.stages >= 2, "wgrad smem budget exceeded");
  p.dW = dW;
  p.trace = trace_slot(); if (p.trace) trace_tag("wgrad k%d %dx%d @%d", ks, Ci, Co, D);
  CUtensorMap mx, mdy, ms;
  memset(&ms, 0, sizeof(ms));
  B200_TRY(make_brick_map(&mx, x + x_coff, x_pitch, Ci, p.atom_ch, N, D, H, W));
  B200_TRY(make_brick_map(&mdy, dy + dy_coff, dy_pitch, Co, p.b_ch, N, D, H, W));
  if (p.scratch) B200_TRY(make_store_map(&ms, p.scratch, (long)p.n_mtiles * 128, Co, Co, p.vsplit, (long)p.n_mtiles * 128 * Co, 4));
  size_t smem = (size_t)p.stages * a_stage + 2 * (size_t)b_bytes + 1024 + 256;
  static bool attr_done = false;
  if (!attr_done) { B200_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr_done = true; }
  B200_CUDA(launch_pdl(wgrad_kernel, dim3(p.n_groups * p.vsplit), dim3(192), smem, st, mx, mdy, ms, p));
  B200_LAUNCH_CHECK();
  if (p.scratch) {
    const int rows = p.n_mtiles * 128;
    const long total = (long)rows * Co;
    B200_CUDA(launch_pdl(wgrad_reduce_kernel, dim3((unsigned)std::min<long>(148L * 8, (total + 255) / 256)), dim3(256), 0, st, (const float*)p.scratch, p.vsplit, rows, Co, Ci,
                         p.taps, p.atom_ch, p.atoms_per_tap, p.total_atoms, dW));
    B200_LAUNCH_CHECK();
  }
  return 0;
}
static void wgrad_plan(WgradParams& p, int Ci, int Co, int ks, int N, int D, int H, int W) {
  p.N = N; p.D = D; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.ks = ks; p.taps = ks * ks * ks;
  p.atom_ch = Ci < 64 ? Ci : 64; p.a_row_bytes = p.atom_ch * 2; p.a_atom_bytes = 128 * p.a_row_bytes;
  p.atoms_per_tap = Ci / p.atom_ch; p.atoms_per_tile = 128 / p.atom_ch; p.total_atoms = p.taps * p.atoms_per_tap;
  p.n_mtiles = cdiv(p.total_atoms, p.atoms_per_tile);
  p.mt_per_cta = 512 / Co; if (p.mt_per_cta > p.n_mtiles) p.mt_per_cta = p.n_mtiles;
  p.n_groups = cdiv(p.n_mtiles, p.mt_per_cta);
  p.mt_per_cta = cdiv(p.n_mtiles, p.n_groups);   // balance the groups
  p.b_ch = Co < 64 ? Co : 64; p.b_row_bytes = p.b_ch * 2; p.b_chunks = Co / p.b_ch; p.b_chunk_bytes = 128 * p.b_row_bytes;
  p.tiles_w = cdiv(W, TW); p.tiles_h = cdiv(H, TH); p.tiles_d = cdiv(D, TD);
  p.total_tiles = (long)N * p.tiles_d * p.tiles_h * p.tiles_w;
  long vs = num_sms() / p.n_groups; if (vs < 1) vs = 1; if (vs > p.total_tiles) vs = p.total_tiles;
  p.vsplit = (int)vs;
  uint32_t cols = (uint32_t)p.mt_per_cta * Co, pw = 32; while (pw < cols) pw <<= 1; p.tmem_cols = pw;
  { uint32_t a_stage = (uint32_t)p.atoms_per_tile * p.a_atom_bytes, b_bytes = (uint32_t)p.b_chunks * p.b_chunk_bytes; p.stages = (int)((200 * 1024 - 2 * b_bytes) / a_stage); if (p.stages > 6) p.stages = 6; }
}

}  // namespace tc
}  // namespace b200
