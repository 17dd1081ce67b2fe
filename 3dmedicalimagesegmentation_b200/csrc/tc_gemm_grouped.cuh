// Grouped tcgen05 GEMM: ONE persistent launch walks a table of independent problems
//
//   D_p[m,n] = sum_k A_p(m,k) * B_p(n,k)        p = 0 .. count-1, fp32 row-major outputs
//
// Used for the ViT weight gradients (dW = dY^T X: both operands MN-major views of [tokens, features] activations).  The
// backward's critical path is the dgrad chain; the 4 weight-gradient GEMMs of a transformer block only feed the optimizer, so the
// executor defers them and issues the GEMMs of several blocks as one launch: 864 equal tiles (4 blocks of ViT-B at 128 x 256) instead of
// 16 launches of 18..72 tiles that each pay a pipeline fill and a ragged last wave on 148 SMs (SURVEY K3/K5-K7 backward,
// /root/reference/unetr.py:69-76 -> 12 x {qkv, out_proj, linear1, linear2}).
//
// Same warp roles as tc::gemm_kernel (tc_gemm.cuh): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue with a
// double-buffered 2 x BN-column accumulator, so the epilogue of tile i (128 x BN fp32 stores) overlaps the MMAs of tile i+1.  The
// per-problem tensor maps live in the __grid_constant__ parameter block; a tile index is mapped to (problem, tm, tn) by a scan of the
// <= 16 tile_begin offsets.
//
// Epilogue = TMA stores.  tcgen05.ld hands every lane one accumulator ROW; stored straight to global memory that is 32 rows x 16 B
// per instruction (measured on B200, 16 problems of 4 ViT-B blocks: 81.8 us with the row-per-lane stores against 38.5 us for the same
// launch with the stores compiled out -- the output leaves at 1.4 TB/s).  Instead each epilogue warp parks 32 rows x 32 columns
// in a 128B-swizzled 4 KB smem box (conflict-free 16-byte st.shared: chunk ^ (row & 7)) and one lane issues
// cp.async.bulk.tensor.2d.global.shared::cta; two boxes per warp alternate so the TMEM reads of chunk c+1 overlap the store of c.
// Ragged M / N edges are clipped by the tensor map.
#pragma once
#include "tc_gemm.cuh"

namespace b200 {
namespace tc {

static constexpr int kMaxGroup = 16;
static constexpr int kGroupStageBytes = 4 * 2 * 4096;

struct GroupProblem {
  float* out; long ldo;          // D row-major [M, N]
  int M, N, K;
  int tiles_m, tiles_n;
  int tile_begin;                // first global tile index of this problem
};

struct alignas(64) GroupParams {
  CUtensorMap map_a[kMaxGroup];
  CUtensorMap map_b[kMaxGroup];
  CUtensorMap map_o[kMaxGroup];  // fp32 output [M, N], box 32 x 32, 128B swizzle
  GroupProblem pr[kMaxGroup];
  int count, total_tiles, BN, stages;
  int dbg_flags;                 // tuning aid (B200_GROUP_DBG): 1 = epilogue reads TMEM but does not store, 2 = row-per-lane st.global epilogue
  uint32_t tmem_cols;
  long long* trace;
};

__device__ __forceinline__ int group_find(const GroupParams& gp, int t) {
  int i = 0;
#pragma unroll 1
  while (i + 1 < gp.count && t >= gp.pr[i + 1].tile_begin) ++i;
  return i;
}

// A_MN / B_MN: operand majors (compile time; every problem of a launch shares them)
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1)
gemm_grouped_kernel(const __grid_constant__ GroupParams gp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = BM * BK * 2, b_bytes = (uint32_t)gp.BN * BK * 2, stage_bytes = a_bytes + b_bytes;
  uint8_t* stg_base = smem + (size_t)gp.stages * stage_bytes;                  // 4 epilogue warps x 2 boxes x 4 KB (1024-byte aligned)
  uint64_t* full = (uint64_t*)(stg_base + kGroupStageBytes);
  uint64_t* empty = full + gp.stages;
  uint64_t* tfull = empty + gp.stages;   // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < gp.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(gp.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  trace_start(gp.trace);

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    int stage = 0; uint32_t phase = 0;
    const uint32_t smem_base_u = smem_u32(smem);
    for (int t = blockIdx.x; t < gp.total_tiles; t += gridDim.x) {
      const int pi = group_find(gp, t);
      const GroupProblem& pr = gp.pr[pi];
      const int r = t - pr.tile_begin;
      const int tm = r % pr.tiles_m, tn = r / pr.tiles_m;
      const int m0 = tm * BM, n0 = tn * gp.BN;
      const int kblocks = (pr.K + BK - 1) / BK;
      const CUtensorMap* ma = &gp.map_a[pi]; const CUtensorMap* mb = &gp.map_b[pi];
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(empty + stage, phase ^ 1);
        if (elect_one()) {
          const uint32_t sa = smem_base_u + (uint32_t)stage * stage_bytes, sb = sa + a_bytes;
          const int k0 = kb * BK;
          mbar_expect_tx(full + stage, stage_bytes);
          if (!A_MN) tma_load_4d(sa, ma, full + stage, k0, m0, 0, 0);
          else { tma_load_4d(sa, ma, full + stage, m0, k0, 0, 0); tma_load_4d(sa + 64 * BK * 2, ma, full + stage, m0 + 64, k0, 0, 0); }
          if (!B_MN) tma_load_4d(sb, mb, full + stage, k0, n0, 0, 0);
          else for (int c = 0; c < gp.BN / 64; ++c) tma_load_4d(sb + c * (64 * BK * 2), mb, full + stage, n0 + c * 64, k0, 0, 0);
        }
        __syncwarp();
        if (++stage == gp.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) |
                           ((uint32_t)(gp.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t hi = desc_hi(1024, 2);
    const uint32_t smem_base_u = smem_u32(smem);
    const uint32_t a_lo0 = desc_lo(smem_base_u, A_MN ? 64 * BK * 2 : 16), b_lo0 = desc_lo(smem_base_u + a_bytes, B_MN ? 64 * BK * 2 : 16);
    const uint32_t a_step = A_MN ? (2048u >> 4) : (32u >> 4), b_step = B_MN ? (2048u >> 4) : (32u >> 4);
    const uint32_t stage_units = stage_bytes >> 4;
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < gp.total_tiles; t += gridDim.x) {
      const int pi = group_find(gp, t);
      const int kblocks = (gp.pr[pi].K + BK - 1) / BK;
      mbar_wait(tempty + acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * gp.BN);
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(full + stage, phase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_units, b_lo = b_lo0 + (uint32_t)stage * stage_units;
        if (elect_one()) {
          umma_f16(tmem_d, desc64(a_lo, hi), desc64(b_lo, hi), idesc, kb ? 1u : 0u);
          umma_f16(tmem_d, desc64(a_lo + a_step, hi), desc64(b_lo + b_step, hi), idesc, 1u);
          umma_f16(tmem_d, desc64(a_lo + 2 * a_step, hi), desc64(b_lo + 2 * b_step, hi), idesc, 1u);
          umma_f16(tmem_d, desc64(a_lo + 3 * a_step, hi), desc64(b_lo + 3 * b_step, hi), idesc, 1u);
          umma_commit(empty + stage);
          if (kb == kblocks - 1) umma_commit(tfull + acc);
        }
        __syncwarp();
        if (++stage == gp.stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4): plain fp32 rows
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0; int sbuf = 0;
    for (int t = blockIdx.x; t < gp.total_tiles; t += gridDim.x) {
      const int pi = group_find(gp, t);
      const GroupProblem& pr = gp.pr[pi];
      const int r = t - pr.tile_begin;
      const int tm = r % pr.tiles_m, tn = r / pr.tiles_m;
      mbar_wait(tfull + acc, acc_phase);
      tc_fence_after();
      const int m = tm * BM + q * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * gp.BN);
      if (gp.dbg_flags & 2) {
        float* orow = pr.out + (long)m * pr.ldo;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(pr.out) & 15) == 0) && (pr.ldo & 3) == 0;
        for (int c0 = 0; c0 < gp.BN; c0 += 32) {
          float v[32];
          tmem_ld16x2(trow + c0, v, true);
          const int n0 = tn * gp.BN + c0;
          if (m < pr.M && n0 < pr.N) {
            const int nv = min(32, pr.N - n0);
            if (nv == 32 && vec_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(orow + n0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j < nv) orow[n0 + j] = v[j];
            }
          }
        }
      } else {
        const CUtensorMap* mo = &gp.map_o[pi];
        const int mrow0 = tm * BM + q * 32;
        for (int c0 = 0; c0 < gp.BN; c0 += 32) {
          float v[32];
          tmem_ld16x2(trow + c0, v, true);
          const int n0 = tn * gp.BN + c0;
          if (mrow0 < pr.M && n0 < pr.N && !(gp.dbg_flags & 1)) {       // warp-uniform
            uint8_t* box = stg_base + ((warp - 2) * 2 + sbuf) * 4096;
            if (lane == 0) bulk_wait_read<1>();                          // the store that last read this box (two chunks ago) is done with it
            __syncwarp();
            uint8_t* rowp = box + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(rowp + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { tma_store_2d(mo, smem_u32(box), n0, mrow0); bulk_commit(); }
            sbuf ^= 1;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_read<0>();      // the staging boxes must outlive the bulk stores that read them
  }
  tc_fence_before();
  __syncthreads();
  trace_end(gp.trace);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(gp.tmem_cols) : "memory");
  }
}

// fp32 row-major [M, N] output as a 2-D tensor map with 32 x 32 boxes (128-byte rows, 128B swizzle) for the TMA-store epilogue
static int make_out_map(CUtensorMap* map, float* out, long M, long N, long ldo) {
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
  B200_CHECK(((uintptr_t)out & 15) == 0 && (ldo * 4) % 16 == 0, "grouped GEMM output must be 16-byte aligned with a 16-byte multiple row pitch (ld=%ld)", ldo);
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ldo * 4};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (output) failed with code %d (dims %ld x %ld, ld %ld)", (int)r, M, N, ldo);
  return 0;
}

// one problem of a grouped launch, host side
struct GroupItem { Operand A, B; float* out; long ldo; int M, N, K; };

// All problems share the operand majors of the first one.  More than kMaxGroup problems are issued as several launches.
static int gemm_grouped(const GroupItem* items, int n, cudaStream_t st) {
  for (int base = 0; base < n; base += kMaxGroup) {
    const int cnt = n - base < kMaxGroup ? n - base : kMaxGroup;
    GroupParams gp; memset(&gp, 0, sizeof(gp));
    const bool a_mn = items[base].A.mn_major(), b_mn = items[base].B.mn_major();
    int maxN = 0; for (int i = 0; i < cnt; ++i) maxN = items[base + i].N > maxN ? items[base + i].N : maxN;
    gp.BN = maxN >= 256 ? 256 : (maxN >= 128 ? 128 : 64);
    if (const char* e = getenv("B200_GROUP_BN")) { int v = atoi(e); if (v == 64 || v == 128 || v == 256) gp.BN = v; }
    const uint32_t stage_bytes = BM * BK * 2 + gp.BN * BK * 2;
    gp.stages = (int)((227 * 1024 - 1024 - 256 - kGroupStageBytes) / stage_bytes); if (gp.stages > 8) gp.stages = 8;
    gp.tmem_cols = (uint32_t)gp.BN * 2; if (gp.tmem_cols < 32) gp.tmem_cols = 32;
    int tiles = 0; double flops = 0;
    for (int i = 0; i < cnt; ++i) {
      const GroupItem& it = items[base + i];
      B200_CHECK(it.M > 0 && it.N > 0 && it.K > 0, "grouped GEMM: empty problem %d", base + i);
      B200_CHECK(it.A.mn_major() == a_mn && it.B.mn_major() == b_mn, "grouped GEMM: the problems of one launch share their operand majors");
      GroupProblem& pr = gp.pr[i];
      pr.out = it.out; pr.ldo = it.ldo; pr.M = it.M; pr.N = it.N; pr.K = it.K;
      pr.tiles_m = cdiv(it.M, BM); pr.tiles_n = cdiv(it.N, gp.BN); pr.tile_begin = tiles;
      tiles += pr.tiles_m * pr.tiles_n;
      flops += 2.0 * it.M * it.N * it.K;
      B200_TRY(make_map(&gp.map_a[i], it.A, it.M, it.K, BM, 1, 1));
      B200_TRY(make_map(&gp.map_b[i], it.B, it.N, it.K, gp.BN, 1, 1));
      B200_TRY(make_out_map(&gp.map_o[i], it.out, it.M, it.N, it.ldo));
    }
    gp.count = cnt; gp.total_tiles = tiles;
    if (const char* e = getenv("B200_GROUP_DBG")) gp.dbg_flags = atoi(e);
    static bool attr_done = false;
    if (!attr_done) {
      B200_CUDA(cudaFuncSetAttribute(gemm_grouped_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      B200_CUDA(cudaFuncSetAttribute(gemm_grouped_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_done = true;
    }
    B200_CHECK(a_mn == b_mn, "grouped GEMM: instantiated for K-major/K-major and MN-major/MN-major operand pairs");
    gp.trace = trace_slot(); if (gp.trace) trace_tag("gemm_grouped x%d tiles %d bn %d gflop %.3f", cnt, tiles, gp.BN, flops * 1e-9);
    const size_t smem = (size_t)gp.stages * stage_bytes + kGroupStageBytes + 1024 + 256;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    cudaError_t le = a_mn ? launch_pdl(gemm_grouped_kernel<true, true>, dim3(grid), dim3(192), smem, st, gp)
                          : launch_pdl(gemm_grouped_kernel<false, false>, dim3(grid), dim3(192), smem, st, gp);
    B200_CUDA(le);
    B200_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace tc

// test hook: n problems, problem i: out_i [M_i, N_i] fp32 = A_i B_i^T with a_i [K_i, M_i], b_i [K_i, N_i] (mn != 0: MN-major) or
// a_i [M_i, K_i], b_i [N_i, K_i] (K-major)
static int tc_gemm_grouped_test(const bf16* const* a, const bf16* const* b, float* const* out, const int* M, const int* N, const int* K, int n, int mn,
                                cudaStream_t st) {
  B200_CHECK(n >= 1 && n <= 64, "1..64 problems");
  tc::GroupItem items[64];
  for (int i = 0; i < n; ++i) {
    items[i].A = mn ? tc::operand(a[i], 1, M[i]) : tc::operand(a[i], K[i], 1);
    items[i].B = mn ? tc::operand(b[i], 1, N[i]) : tc::operand(b[i], K[i], 1);
    items[i].out = out[i]; items[i].ldo = N[i]; items[i].M = M[i]; items[i].N = N[i]; items[i].K = K[i];
  }
  return tc::gemm_grouped(items, n, st);
}

}  // namespace b200
