// tcgen05 / TMEM / TMA GEMM engine for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
//   D[b][m,n] = sum_k A[b](m,k) * B[b](n,k)          b = (b0,b1) batch pair, epilogue functor decides the store
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   (cp.async.bulk.tensor.4d -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      MMA issuer     (one elected lane: tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16 per instruction;
//                               tcgen05.commit frees smem stages / publishes the accumulator); also owns TMEM alloc
//   warps 2..5  epilogue       (tcgen05.ld 32x32b.x16 -> registers -> functor; TMEM accumulator double-buffered so the
//                               epilogue of tile i overlaps the MMAs of tile i+1)
// Either operand may be K-major (row = m/n, K contiguous) or MN-major (row = k, m/n contiguous): linear-layer
// dgrad/wgrad and the attention P.V / dS^T.Q products need no transposed copies.  Ragged M/N/K edges are handled
// by TMA out-of-bounds zero fill plus an m<M, n<N guard in the epilogue.
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <utility>

#include "common.cuh"
#include "contract.cuh"

namespace b200 {
namespace tc {

static constexpr int BM = 128, BK = 64;
static constexpr int EPI_STG_BYTES = 4 * 32 * 36 * 4;   // coalesced-epilogue staging: 4 warps x 32 rows x (32 + 4 pad) floats

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Straight-line asm (internal retry loop) so the compiler keeps the surrounding loops warp-uniform.
// Watchdog: ~4M failed probes (each probe suspends in hardware for a while) -> trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .u32 n;\n"
      "mov.u32 n, 0;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "add.u32 n, n, 1;\n"
      "setp.gt.u32 p, n, 4194304;\n"
      "@p trap;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// one elected lane of a converged warp (keeps the surrounding loop warp-uniform so operands stay in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
// descriptor halves: lo = start>>4 | (LBO>>4)<<16 ; hi = SBO>>4 | version(1)<<14 | layout<<29
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo_bytes) { return ((addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 consecutive TMEM columns of this thread's lane as two x16 loads behind ONE wait (the second is skipped when `two` is false;
// its registers are then zero)
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr, float* v, bool two) {
  uint32_t r[32];
#pragma unroll
  for (int i = 16; i < 32; ++i) r[i] = 0u;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  if (two)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr + 16u) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) { asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory"); }

// ---- TMA-store epilogue helpers (bulk tensor stores global <- shared::cta, tracked by bulk async-groups of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// one accumulator row of 32 columns into a [32 rows][32 cols] smem box in the layout the tensor map expects: fp32 rows are 128 B
// (SWIZZLE_128B: 16-byte chunk ^ (row & 7)), bf16 rows 64 B (SWIZZLE_64B: chunk ^ ((row >> 1) & 3)); both conflict-free per quarter-warp
__device__ __forceinline__ void stage_row32(uint8_t* box, int row, const float* v, float) {
  uint8_t* rp = box + row * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(rp + ((j ^ (row & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void stage_row32(uint8_t* box, int row, const float* v, bf16) {
  uint8_t* rp = box + row * 64;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
    *reinterpret_cast<uint4*>(rp + ((j ^ ((row >> 1) & 3)) << 4)) = t;
  }
}
// epilogue functors that can hand their values to the TMA-store path (EpStore<TO>, contract.cuh)
template <class EP> struct ep_tma { static constexpr bool ok = false; typedef float out_t; };
template <class TO> struct ep_tma<EpStore<TO>> { static constexpr bool ok = true; typedef TO out_t; };

// shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=2 [61,64))
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// 16 consecutive output columns of one row: functors with a vectorised `seg16` use it, others get a rolled scalar loop
template <class EP>
__device__ __forceinline__ auto ep_seg16(const EP& ep, int b, int m, int n0, const float* v, int nvalid) -> decltype(ep.seg16(b, m, n0, v, nvalid), void()) {
  ep.seg16(b, m, n0, v, nvalid);
}
template <class EP, class... Dummy>
__device__ __forceinline__ void ep_seg16(const EP& ep, int b, int m, int n0, const float* v, int nvalid, Dummy...) {
#pragma unroll 1
  for (int j = 0; j < nvalid; ++j) ep(b, m, n0 + j, v[j]);
}

// 4 consecutive output columns of one row (coalesced epilogue): functors with `seg4` use it, others get the scalar call
template <class EP>
__device__ __forceinline__ auto ep_seg4(const EP& ep, int b, int m, int n0, const float* v, int nvalid) -> decltype(ep.seg4(b, m, n0, v, nvalid), void()) {
  ep.seg4(b, m, n0, v, nvalid);
}
template <class EP, class... Dummy>
__device__ __forceinline__ void ep_seg4(const EP& ep, int b, int m, int n0, const float* v, int nvalid, Dummy...) {
#pragma unroll 1
  for (int j = 0; j < nvalid; ++j) ep(b, m, n0 + j, v[j]);
}

// functors that provide seg4 take the coalesced epilogue; the others (atomic / scatter epilogues) keep one row per lane
template <class EP> struct has_seg4 {
  template <class U> static constexpr auto test(int) -> decltype(std::declval<const U&>().seg4(0, 0, 0, (const float*)nullptr, 0), true) { return true; }
  template <class U> static constexpr bool test(...) { return false; }
  static constexpr bool value = test<EP>(0);
};

// Row-wise epilogues (EP::kRowOp == 1 forward softmax, 2 softmax backward): the CTA's tile holds COMPLETE rows (tiles_n == 1), each
// epilogue thread owns one row and walks its TMEM columns twice -- attention scores never go to memory as fp32.
template <class EP> struct row_op_of { template <class U> static constexpr int get(decltype(U::kRowOp)*) { return U::kRowOp; }
                                       template <class U> static constexpr int get(...) { return 0; }
                                       static constexpr int value = get<EP>(nullptr); };

struct Params {
  int M, N, K, BN;          // BN in {64,128,256}
  int a_mn, b_mn;           // operand majors
  int tiles_m, tiles_n, nb1, batches;
  int stages;
  uint32_t tmem_cols;       // 2*BN rounded to a power of two
  int ksplit, kb_per_split; // split-K: work item = (tile, split); epilogue functor must accumulate atomically
  int coalesce;             // epilogue stores through a per-warp smem transpose (8 lanes x 16 B per row) instead of one row per lane
  int tma_store;            // epilogue through swizzled smem boxes + cp.async.bulk.tensor stores (plain dense EpStore outputs); 2 = with second output
  uint32_t stg_bytes;       // staging area of the TMA-store epilogue (between the stage ring and the barriers)
  long long* dbg;           // optional: CTA 0 phase timestamps (clock64) for tuning
  long long* trace;         // optional in-situ (start, end) slot
};

template <class EP, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_o,
            const __grid_constant__ CUtensorMap map_o2, const Params p, const EP ep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = BM * BK * 2, b_bytes = (uint32_t)p.BN * BK * 2, stage_bytes = a_bytes + b_bytes;
  uint8_t* stg_base = smem + (size_t)p.stages * stage_bytes;                 // TMA-store boxes (1024-byte aligned), p.stg_bytes
  uint64_t* full = (uint64_t*)(stg_base + p.stg_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tfull = empty + p.stages;   // [2]
  uint64_t* tempty = tfull + 2;         // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
#ifdef B200_COALESCED_EPILOGUE
  float* stg_all = (float*)(stg_base + p.stg_bytes + 256);   // 4 epilogue warps x 32 rows x 36 floats (EPI_STG_BYTES)
#endif

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dbg = p.dbg && blockIdx.x == 0;
  if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();
  if (p.dbg && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); p.dbg[64 + 2 * blockIdx.x] = g; }
  const int kblocks_all = (p.K + BK - 1) / BK;
  const long total_tiles = (long)p.tiles_m * p.tiles_n * p.batches * p.ksplit;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a); prefetch_tmap(&map_b);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 4); }
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
  pdl_wait();   // predecessor's global writes are visible from here on
  trace_start(p.trace);
  if (dbg && threadIdx.x == 0) p.dbg[1] = clock64();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (warp-uniform loop, one elected lane issues)
    int stage = 0; uint32_t phase = 0;
    const uint32_t smem_base_u = smem_u32(smem);
    for (long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int sp = (int)(t % p.ksplit); long r = t / p.ksplit;
      int tm = (int)(r % p.tiles_m); r /= p.tiles_m; int tn = (int)(r % p.tiles_n); int b = (int)(r / p.tiles_n);
      int b0 = b / p.nb1, b1 = b % p.nb1;
      const int kb_begin = sp * p.kb_per_split, kb_end = min(kblocks_all, kb_begin + p.kb_per_split);
      const int m0 = tm * BM, n0 = tn * p.BN;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(empty + stage, phase ^ 1);
        if (elect_one()) {
          const uint32_t sa = smem_base_u + (uint32_t)stage * stage_bytes, sb = sa + a_bytes;
          const int k0 = kb * BK;
          mbar_expect_tx(full + stage, stage_bytes);
          if (!A_MN) tma_load_4d(sa, &map_a, full + stage, k0, m0, b1, b0);
          else { tma_load_4d(sa, &map_a, full + stage, m0, k0, b1, b0); tma_load_4d(sa + 64 * BK * 2, &map_a, full + stage, m0 + 64, k0, b1, b0); }
          if (!B_MN) tma_load_4d(sb, &map_b, full + stage, k0, n0, b1, b0);
          else for (int c = 0; c < p.BN / 64; ++c) tma_load_4d(sb + c * (64 * BK * 2), &map_b, full + stage, n0 + c * 64, k0, b1, b0);
          if (dbg && kb < 4 && t == blockIdx.x) p.dbg[16 + kb] = clock64();
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    if (dbg && lane == 0) p.dbg[2] = clock64();   // producer done issuing
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues)
    // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=BF16 [7,10), b=BF16 [10,13),
    // a_major bit15, b_major bit16, N>>3 [17,23), M>>4 [24,29)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)A_MN << 15) | ((uint32_t)B_MN << 16) |
                           ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    // K-major: 8-row groups 1024 B apart, +32 B per 16-wide K step inside the swizzle atom.
    // MN-major: 64-wide MN atoms 64*BK*2 B apart (LBO), 8-k-row groups 1024 B apart (SBO), +2048 B per K step.
    const uint32_t hi = desc_hi(1024, 2);
    const uint32_t smem_base_u = smem_u32(smem);
    const uint32_t a_lo0 = desc_lo(smem_base_u, A_MN ? 64 * BK * 2 : 16), b_lo0 = desc_lo(smem_base_u + a_bytes, B_MN ? 64 * BK * 2 : 16);
    const uint32_t a_step = A_MN ? (2048u >> 4) : (32u >> 4), b_step = B_MN ? (2048u >> 4) : (32u >> 4);
    const uint32_t stage_units = stage_bytes >> 4;
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
    for (long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int sp = (int)(t % p.ksplit);
      const int kblocks = min(kblocks_all, (sp + 1) * p.kb_per_split) - sp * p.kb_per_split;
      mbar_wait(tempty + acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.BN);
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(full + stage, phase);
        tc_fence_after();
        if (dbg && lane == 0 && kb < 4 && t == blockIdx.x) p.dbg[8 + kb] = clock64();   // arrival of the first k-blocks
        const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_units, b_lo = b_lo0 + (uint32_t)stage * stage_units;
        if (elect_one()) {
          umma_f16(tmem_d, desc64(a_lo, hi), desc64(b_lo, hi), idesc, kb ? 1u : 0u);
          umma_f16(tmem_d, desc64(a_lo + a_step, hi), desc64(b_lo + b_step, hi), idesc, 1u);
          umma_f16(tmem_d, desc64(a_lo + 2 * a_step, hi), desc64(b_lo + 2 * b_step, hi), idesc, 1u);
          umma_f16(tmem_d, desc64(a_lo + 3 * a_step, hi), desc64(b_lo + 3 * b_step, hi), idesc, 1u);
          umma_commit(empty + stage);                 // smem stage reusable once these MMAs retire
          if (kb == kblocks - 1) umma_commit(tfull + acc);  // accumulator complete
          if (dbg && kb < 4 && t == blockIdx.x) p.dbg[12 + kb] = clock64();   // MMAs of this k-block issued
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (dbg && lane == 0) p.dbg[3] = clock64();   // MMA issue done
  } else {
    // ------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0; int tsb = 0;
    for (long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      long r = t / p.ksplit;
      int tm = (int)(r % p.tiles_m); r /= p.tiles_m; int tn = (int)(r % p.tiles_n); int b = (int)(r / p.tiles_n);
      if (p.ksplit > 1) b += (int)(t % p.ksplit) * p.batches;   // split-K functors decode (split, batch) = (b / batches, b % batches)
      mbar_wait(tfull + acc, acc_phase);
      tc_fence_after();
      if (dbg && warp == 2 && lane == 0 && t == blockIdx.x) p.dbg[4] = clock64();   // first accumulator ready
      const int m = tm * BM + q * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
      if constexpr (row_op_of<EP>::value == 1) {
        // P = softmax(scale * S) over the N valid columns; two passes over TMEM (running max / sum, then normalise + store)
        float mx = -INFINITY, sum = 0.f;
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          float v[16];
          tmem_ld16(trow + c0, v);
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = (c0 + j < p.N) ? v[j] * ep.scale : -INFINITY; cm = fmaxf(cm, v[j]); }
          const float nm = fmaxf(mx, cm);
          float cs = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) cs += __expf(v[j] - nm);
          sum = sum * __expf(mx - nm) + cs; mx = nm;
        }
        const float inv = 1.f / sum;
        for (int c0 = 0; c0 < ep.ldw; c0 += 16) {      // ldw >= N: padding columns are written as zeros
          float v[16];
          if (c0 < p.N) tmem_ld16(trow + c0, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (c0 + j < p.N) ? __expf(v[j] * ep.scale - mx) * inv : 0.f;
          if (m < p.M) ep.store16(b, m, c0, v, min(16, ep.ldw - c0));
        }
      } else if constexpr (row_op_of<EP>::value == 2) {
        // dS = P * (dP - sum_j dP_j P_j) * scale, P read back from memory (bf16), dP = this accumulator row
        float dot = 0.f;
        const bool rv = m < p.M;               // every lane runs the (warp-aligned) TMEM loads; only valid rows touch memory
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          float v[16], pr[16];
          tmem_ld16(trow + c0, v);
          ep.load_p16(b, m, c0, pr, rv ? min(16, p.N - c0) : 0);
#pragma unroll
          for (int j = 0; j < 16; ++j) dot = fmaf(v[j], pr[j], dot);
        }
        for (int c0 = 0; c0 < ep.ldw; c0 += 16) {
          float v[16], pr[16];
          if (c0 < p.N) tmem_ld16(trow + c0, v);
          ep.load_p16(b, m, c0, pr, (rv && c0 < p.N) ? min(16, p.N - c0) : 0);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (c0 + j < p.N) ? pr[j] * (v[j] - dot) * ep.scale : 0.f;
          if (rv) ep.store16(b, m, c0, v, min(16, ep.ldw - c0));
        }
      } else {
        bool done = false;
        if constexpr (ep_tma<EP>::ok) {
          if (p.tma_store) {
            // Row-per-lane st.global is what bounds the epilogue of these single-tile GEMMs (measured: 432x3072x768, 128 CTAs: MMAs
            // done at cycle 9.5 k, kernel end at 17.3 k -- 4 us of scattered 16-byte stores).  Values go through swizzled smem boxes
            // of 32 rows x 32 columns instead and ONE lane hands each box to the copy engine; two boxes per output alternate.
            typedef typename ep_tma<EP>::out_t TO;
            constexpr uint32_t BOX = 32 * 32 * sizeof(TO);
            const bool two = p.tma_store == 2;
            uint8_t* wstg = stg_base + (uint32_t)(warp - 2) * (two ? 4u : 2u) * BOX;
            const int mrow0 = tm * BM + q * 32;
            const int z = p.ksplit > 1 ? (int)(t % p.ksplit) : 0;
            for (int c0 = 0; c0 < p.BN; c0 += 32) {
              float a32[32];
              tmem_ld16x2(trow + c0, a32, true);
              const int n0 = tn * p.BN + c0;
              if (mrow0 < p.M && n0 < p.N) {                 // warp-uniform
                float v[32], pv[32];
                ep.values16(m < p.M, m, n0, a32, v, pv);
                ep.values16(m < p.M, m, n0 + 16, a32 + 16, v + 16, pv + 16);
                uint8_t* box = wstg + (uint32_t)tsb * BOX;
                if (lane == 0) bulk_wait_read<1>();           // the group that last read these boxes (two chunks ago) is done
                __syncwarp();
                stage_row32(box, lane, v, TO());
                if (two) stage_row32(box + 2 * BOX, lane, pv, TO());
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                  tma_store_3d(&map_o, smem_u32(box), n0, mrow0, z);
                  if (two) tma_store_3d(&map_o2, smem_u32(box + 2 * BOX), n0, mrow0, 0);
                  bulk_commit();
                }
                tsb ^= 1;
              }
            }
            done = true;
          }
        }
#ifdef B200_COALESCED_EPILOGUE
        // EXPERIMENT, compiled out (make EXTRA=-DB200_COALESCED_EPILOGUE): measured SLOWER on B200 -- 7.44 vs 6.73 ms/step; in situ the
        // single-tile GEMMs grow by ~5 us each (432x3072x768: 14.1 -> 19.8 us, with the GELU-backward epilogue 19.9 -> 34.9 us).  The
        // row-per-lane stores are not what bounds the epilogue: the 16 smem operations + 8 address computations per 32 columns on a
        // warp that is alone on its SM sub-partition cost more than the sector savings return.
        if constexpr (has_seg4<EP>::value) {
          if (p.coalesce) {
            // One row per lane is what tcgen05.ld delivers, and stored that way every warp store touches 32 rows (32 sectors per
            // instruction, ~32 LSU cycles each).  32 columns at a time go through a per-warp smem transpose (row pitch 36 floats:
            // conflict-free both ways) so that 8 lanes cover 32 consecutive columns of one row.
            float* stg = stg_all + (warp - 2) * (32 * 36);
            const auto tl = ep.tile(b);
            const int ch = lane & 7, rsub = lane >> 3;
            for (int c0 = 0; c0 < p.BN; c0 += 32) {
              float v[32];
              tmem_ld16x2(trow + c0, v, c0 + 16 < p.BN);
#pragma unroll
              for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(stg + lane * 36 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              __syncwarp();
              const int cc = c0 + 4 * ch;
              const int n0 = tn * p.BN + cc;
              const int nv = min(4, p.N - n0);
              if (cc < p.BN && nv > 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int r = i * 4 + rsub;
                  const float4 t = *reinterpret_cast<const float4*>(stg + r * 36 + 4 * ch);
                  const int mm = tm * BM + q * 32 + r;
                  if (mm < p.M) ep.seg4t(tl, b, mm, n0, reinterpret_cast<const float*>(&t), nv);
                }
              }
              __syncwarp();
            }
            done = true;
          }
        }
#endif
        if (!done) {
          for (int c0 = 0; c0 < p.BN; c0 += 16) {
            float v[16];
            tmem_ld16(trow + c0, v);
            int n0 = tn * p.BN + c0;
            if (m < p.M && n0 < p.N) ep_seg16(ep, b, m, n0, v, min(16, p.N - n0));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_store && lane == 0) bulk_wait_read<0>();      // the staging boxes must outlive the bulk stores that read them
  }
  tc_fence_before();
  __syncthreads();
  trace_end(p.trace);
  if (dbg && threadIdx.x == 0) p.dbg[5] = clock64();
  if (p.dbg && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); p.dbg[65 + 2 * blockIdx.x] = g; }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

// A bf16 operand: element (o, k, b1, b0) at p[o*so + k*sk + b1*sb1 + b0*sb0] where exactly one of so/sk is 1.
struct Operand {
  const bf16* p; long so, sk, sb0, sb1;
  bool mn_major() const { return so == 1 && sk != 1; }
};
static inline Operand operand(const bf16* p, long so, long sk, long sb0 = 0, long sb1 = 0) { Operand o{p, so, sk, sb0, sb1}; return o; }

static int make_map(CUtensorMap* map, const Operand& op, long extent_o, long extent_k, int box_o, int nb0, int nb1) {
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
  bool mn = op.mn_major();
  long ld = mn ? op.sk : op.so;
  B200_CHECK(((uintptr_t)op.p & 15) == 0 && (ld * 2) % 16 == 0, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ld=%ld)", ld);
  B200_CHECK((nb1 <= 1 || (op.sb1 * 2) % 16 == 0) && (nb0 <= 1 || (op.sb0 * 2) % 16 == 0), "TMA batch strides must be 16-byte multiples");
  cuuint64_t dims[4] = {(cuuint64_t)(mn ? extent_o : extent_k), (cuuint64_t)(mn ? extent_k : extent_o), (cuuint64_t)nb1, (cuuint64_t)nb0};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)(nb1 > 1 ? op.sb1 * 2 : ld * 2), (cuuint64_t)(nb0 > 1 ? op.sb0 * 2 : ld * 2)};
  cuuint32_t box[4] = {64u, (cuuint32_t)(mn ? BK : box_o), 1u, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)op.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (dims %lu x %lu, ld %ld)", (int)r, (unsigned long)dims[0], (unsigned long)dims[1], ld);
  return 0;
}

// row-major [Z][M][N] output of element size `esz` (4: fp32, 2: bf16) as a 3-D tensor map with 32 x 32 boxes for the TMA-store epilogue
static int make_store_map(CUtensorMap* map, void* out, long M, long N, long ld, long Z, long zstride, int esz) {
  EncodeTiledFn enc = get_encode();
  B200_CHECK(enc, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)(Z > 1 ? Z : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * esz, (cuuint64_t)(Z > 1 ? zstride : ld * M) * esz};
  cuuint32_t box[3] = {32u, 32u, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, esz == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (store map) failed with code %d (dims %ld x %ld x %ld, ld %ld)", (int)r, N, M, Z, ld);
  return 0;
}
template <class EP> static inline bool ep_tma_ok(const EP&, int) { return false; }
template <class TO> static inline bool ep_tma_ok(const EpStore<TO>& ep, int nbatch) { return ep.tma_store_ok(nbatch); }
template <class EP> static inline int ep_make_store_maps(const EP&, CUtensorMap*, CUtensorMap*, int, int, int) { return 0; }
template <class TO> static inline int ep_make_store_maps(const EpStore<TO>& ep, CUtensorMap* mo, CUtensorMap* mo2, int M, int N, int ksplit) {
  B200_TRY(make_store_map(mo, (void*)ep.out, M, N, ep.ld, ep.splitk_nbat ? ksplit : 1, ep.split_stride, (int)sizeof(TO)));
  if (ep.preact) B200_TRY(make_store_map(mo2, (void*)ep.preact, M, N, ep.ld, 1, 0, (int)sizeof(TO)));
  return 0;
}
template <class EP> static inline bool ep_has_second(const EP&) { return false; }
template <class TO> static inline bool ep_has_second(const EpStore<TO>& ep) { return ep.preact != nullptr; }

static long long* g_dbg = nullptr;   // set by b200_test_set_debug_buffer
static int g_num_sms = 0;
static inline int num_sms() {
  if (!g_num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev); if (g_num_sms <= 0) g_num_sms = 148; }
  return g_num_sms;
}

// split-K factor worth using for a fp32-atomic epilogue: few output tiles, long K
static inline int plan_splitk(int M, int N, int K, int nb) {
  // (fp32 atomics in the epilogue were measured slower than the K loop they save: proj 10.6 -> 17.6 us; the partial tiles are
  // stored plainly and summed by the consumer kernel -- see SplitSum in elementwise.cuh)
  static const bool off = getenv("B200_NO_SPLITK") != nullptr;
  if (off) return 1;
  const long tiles = (long)cdiv(M, BM) * cdiv(N, 64) * nb; const int kbs = cdiv(K, BK);
  if (tiles * 2 > num_sms() || kbs < 8) return 1;
  long s = num_sms() / tiles; if (s > kbs / 4) s = kbs / 4; if (s > 4) s = 4;
  return s < 2 ? 1 : (int)s;
}

// D[(b0,b1)][m,n] = sum_k A(m,k) B(n,k);  ep(b0*nb1+b1, m, n, acc)
template <class EP>
static int gemm(const Operand& A, const Operand& B, const EP& ep, int M, int N, int K, int nb0, int nb1, cudaStream_t st, bool allow_splitk = false,
                int ksplit_req = 0) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  Params p;
  p.M = M; p.N = N; p.K = K;
  p.a_mn = A.mn_major(); p.b_mn = B.mn_major();
  // N-tile: minimise (waves) x (per-tile cost); per-tile cost ~ k-blocks x rows fetched (L2->SM bound, measured ~2.2 ns per
  // row of 128 B per k-block in situ) + a fixed pipeline fill + the epilogue.  MN-major B is fetched in 64-column boxes.
  // TMA-store epilogue: plain dense EpStore outputs (one batch, no accumulate, no atomic split-K), 32-column boxes
  static const bool tma_off = getenv("B200_NO_TMA_STORE") != nullptr;
  const bool tma_ok = ep_tma<EP>::ok && !tma_off && !allow_splitk && N % 32 == 0 && ep_tma_ok(ep, nb0 * nb1);
  {
    static const int cand_k[] = {32, 48, 64, 80, 96, 112, 128, 160, 192, 224, 256}, cand_mn[] = {64, 128, 192, 256};
    static const int cand_k32[] = {32, 64, 96, 128, 160, 192, 224, 256};
    const int* cand = p.b_mn ? cand_mn : (tma_ok ? cand_k32 : cand_k); const int nc = p.b_mn ? 4 : (tma_ok ? 8 : 11);
    const long tm = cdiv(M, BM), nb = (long)nb0 * nb1; const int kbs = cdiv(K, BK);
    double best = 1e30; p.BN = 64;
    for (int i = 0; i < nc; ++i) {
      const int bn = cand[i];
      if (bn > 64 && bn - 16 >= N) continue;
      const long tiles = tm * cdiv(N, bn) * nb * (ksplit_req > 1 ? ksplit_req : 1);
      const long waves = (tiles + num_sms() - 1) / num_sms();
      const int kb = ksplit_req > 1 ? cdiv(kbs, ksplit_req) : kbs;
      const double cost = waves * (kb * (128.0 + bn) * 2.2e-3 + 0.6 + bn * 6e-3);
      if (cost < best - 1e-9) { best = cost; p.BN = bn; }
    }
    if (row_op_of<EP>::value) {          // complete rows per tile
      p.BN = (N + 15) & ~15; if (p.BN < 32) p.BN = 32;
      B200_CHECK(!p.b_mn && p.BN <= 256, "row-wise epilogue needs a K-major B operand and N <= 256 (N=%d)", N);
    }
  }
  p.tiles_m = cdiv(M, BM); p.tiles_n = cdiv(N, p.BN); p.nb1 = nb1; p.batches = nb0 * nb1;
  uint32_t stage_bytes = BM * BK * 2 + p.BN * BK * 2;
  p.tma_store = tma_ok ? (ep_has_second(ep) ? 2 : 1) : 0;
  p.stg_bytes = tma_ok ? 4u * (p.tma_store == 2 ? 4u : 2u) * 32u * 32u * (uint32_t)sizeof(typename ep_tma<EP>::out_t) : 0u;
  { const uint32_t avail = 227u * 1024 - 1280 - p.stg_bytes, budget = avail < 200u * 1024 ? avail : 200u * 1024;
    p.stages = (int)(budget / stage_bytes); if (p.stages > 8) p.stages = 8; }
  { uint32_t c = 32; while (c < (uint32_t)p.BN * 2) c <<= 1; p.tmem_cols = c; }
  CUtensorMap ma, mb;
  B200_TRY(make_map(&ma, A, M, K, BM, nb0, nb1));
  B200_TRY(make_map(&mb, B, N, K, p.BN, nb0, nb1));
  static bool attr_done = false;  // per EP instantiation
  if (!attr_done) {
    B200_CUDA(cudaFuncSetAttribute(gemm_kernel<EP, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CUDA(cudaFuncSetAttribute(gemm_kernel<EP, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CUDA(cudaFuncSetAttribute(gemm_kernel<EP, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CUDA(cudaFuncSetAttribute(gemm_kernel<EP, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  p.dbg = g_dbg;
  p.trace = trace_slot(); if (p.trace) trace_tag("gemm %dx%dx%d b%d %s%s", M, N, K, nb0 * nb1, A.mn_major() ? "m" : "k", B.mn_major() ? "m" : "k");
  if (const char* e = getenv("B200_GEMM_BN")) { int v = atoi(e); if (v == 64 || v == 128 || v == 256) { p.BN = v; p.tiles_n = cdiv(N, p.BN); stage_bytes = BM * BK * 2 + p.BN * BK * 2; p.stages = (int)((200 * 1024) / stage_bytes); if (p.stages > 8) p.stages = 8; p.tmem_cols = p.BN * 2; } }
  if (const char* e = getenv("B200_GEMM_STAGES")) { int v = atoi(e); if (v >= 1 && v <= p.stages) p.stages = v; }
  p.ksplit = 1; p.kb_per_split = cdiv(K, BK);
  if (ksplit_req > 1) { int kbs = cdiv(K, BK); p.kb_per_split = cdiv(kbs, ksplit_req); p.ksplit = cdiv(kbs, p.kb_per_split); }
  else if (allow_splitk) {
    long base_tiles = (long)p.tiles_m * p.tiles_n * p.batches; int kbs = cdiv(K, BK);
    long want = num_sms() / base_tiles; if (want > kbs / 4) want = kbs / 4; if (want < 1) want = 1;
    p.kb_per_split = cdiv(kbs, (int)want); p.ksplit = cdiv(kbs, p.kb_per_split);
  }
#ifdef B200_COALESCED_EPILOGUE
  static const bool no_coalesce = getenv("B200_NO_COALESCED_EPILOGUE") != nullptr;
  p.coalesce = no_coalesce ? 0 : 1;
  size_t smem = (size_t)p.stages * stage_bytes + p.stg_bytes + 1024 + 256 + EPI_STG_BYTES;
#else
  p.coalesce = 0;
  size_t smem = (size_t)p.stages * stage_bytes + p.stg_bytes + 1024 + 256;
#endif
  B200_CHECK(smem <= 227 * 1024, "tcgen05 GEMM smem budget exceeded (%zu)", smem);
  CUtensorMap mo, mo2;
  memset(&mo, 0, sizeof(mo)); memset(&mo2, 0, sizeof(mo2));
  if (p.tma_store) B200_TRY(ep_make_store_maps(ep, &mo, &mo2, M, N, p.ksplit));
  long tiles = (long)p.tiles_m * p.tiles_n * p.batches * p.ksplit;
  int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  cudaError_t le;
  if (p.a_mn) { if (p.b_mn) le = launch_pdl(gemm_kernel<EP, true, true>, dim3(grid), dim3(192), smem, st, ma, mb, mo, mo2, p, ep); else le = launch_pdl(gemm_kernel<EP, true, false>, dim3(grid), dim3(192), smem, st, ma, mb, mo, mo2, p, ep); }
  else { if (p.b_mn) le = launch_pdl(gemm_kernel<EP, false, true>, dim3(grid), dim3(192), smem, st, ma, mb, mo, mo2, p, ep); else le = launch_pdl(gemm_kernel<EP, false, false>, dim3(grid), dim3(192), smem, st, ma, mb, mo, mo2, p, ep); }
  B200_CUDA(le);
  B200_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc

// test hook: fp32 out = A B^T with selectable majors.  a: [M,K] (K-major) or [K,M] (MN-major); b likewise.
static int tc_gemm_test(const bf16* a, const bf16* b, float* out, int M, int N, int K, int a_mn, int b_mn, cudaStream_t st) {
  tc::Operand A = a_mn ? tc::operand(a, 1, M) : tc::operand(a, K, 1);
  tc::Operand B = b_mn ? tc::operand(b, 1, N) : tc::operand(b, K, 1);
  return tc::gemm(A, B, ep_plain<float>(out, N), M, N, K, 1, 1, st);
}

}  // namespace b200
