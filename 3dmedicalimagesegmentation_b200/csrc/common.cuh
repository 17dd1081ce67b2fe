// Shared device/host helpers for the B200 UNETR library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <utility>

namespace b200 {

typedef __nv_bfloat16 bf16;

// ---- error plumbing: every C-ABI entry returns 0 or a code; text via b200_last_error() ----
void set_error(const char* fmt, ...);
const char* get_error();

#define B200_CHECK(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      b200::set_error(__VA_ARGS__);      \
      return 1;                          \
    }                                    \
  } while (0)

#define B200_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      b200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                      \
    }                                                                                \
  } while (0)

#define B200_LAUNCH_CHECK()        \
  do {                             \
    ++b200::g_launches;            \
    B200_CUDA(cudaGetLastError()); \
  } while (0)

#define B200_TRY(...)        \
  do {                       \
    int _rc = (__VA_ARGS__); \
    if (_rc) return _rc;     \
  } while (0)

// ---- launch counter + optional per-op CUDA-event profiler (b200_prof_* in the C ABI) ----
extern unsigned long long g_launches;
void prof_begin(const char* tag, cudaStream_t st);
void prof_end(cudaStream_t st);
extern bool g_prof_on;
extern bool g_prof_coarse;   // phase-level regions only (no per-op events: their ~8 us floor hides small kernels)
struct ProfScope {
  cudaStream_t st; bool active;
  ProfScope(const char* tag, cudaStream_t s) : st(s), active(g_prof_on) { if (active) prof_begin(tag, st); }
  ~ProfScope() { if (active) prof_end(st); }
};
#define B200_PROFC_BEGIN(tag, st) do { if (b200::g_prof_coarse) b200::prof_begin(tag, st); } while (0)
#define B200_PROFC_END(st) do { if (b200::g_prof_coarse) b200::prof_end(st); } while (0)
#define B200_PROF(tag, st) b200::ProfScope _prof_scope(tag, st)
// detailed variant: printf-style tag, formatted only when profiling is on
#define B200_PROFD(st, ...)                                   \
  char _prof_tag[96];                                         \
  if (b200::g_prof_on) snprintf(_prof_tag, sizeof(_prof_tag), __VA_ARGS__); \
  b200::ProfScope _prof_scope(_prof_tag, st)

// ---- in-situ kernel trace (b200_trace_* in the C ABI): every tcgen05 launch gets a slot; CTAs record min start / max end of
// %globaltimer, so durations and gaps are measured inside a real step without per-op events
extern long long* g_trace;     // device buffer of (start, end) pairs, or null
extern int g_trace_n, g_trace_cap;
void trace_tag(const char* fmt, ...);
static inline long long* trace_slot() { return (g_trace && g_trace_n < g_trace_cap) ? g_trace + 2 * (g_trace_n++) : nullptr; }
__device__ __forceinline__ void trace_start(long long* slot) {
  if (slot && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); atomicMin(slot, g); }
}
__device__ __forceinline__ void trace_end(long long* slot) {
  if (slot && threadIdx.x == 0) { long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); atomicMax(slot + 1, g); }
}

// ---- programmatic dependent launch (PDL): a kernel launched through launch_pdl may start while its predecessor in the stream is
// still draining; EVERY such kernel calls pdl_wait() before its first global-memory access (reads AND writes: the predecessor may
// still be reading a buffer this kernel overwrites).  What precedes pdl_wait() -- barrier init, TMEM allocation, descriptor
// prefetch, or just the launch latency of a small kernel -- overlaps the predecessor's tail.  B200_NO_PDL=1 turns the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool off = getenv("B200_NO_PDL") != nullptr;
  cfg.attrs = at; cfg.numAttrs = off ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- scalar conversions ----
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(double v) { return (float)v; }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
template <class T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- 16-byte vectors of T viewed as floats ----
template <class T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec16<bf16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const bf16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ void store(bf16* p) const {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

template <> struct Vec16<__half> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const __half* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ __forceinline__ void store(__half* p) const {
    uint4 t;
    __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// Storage type of RAW convolution outputs (the tensors an InstanceNorm consumes).  In bf16 mode they are kept as fp16: same bytes,
// 8x finer rounding.  Measured on the fp32 oracle (tests/test_oracle.py): bf16 rounding of these tensors alone costs 0.65e-2 of the
// 1e-2 logits budget at configs[0] (the normalisation amplifies |mean|/std), fp16 makes that term vanish (total 0.98e-2 -> 0.70e-2).
// Range is safe: pre-norm outputs of an InstanceNorm'd network are O(1..100) against fp16's 65504.
template <class T> struct RawOf { typedef T type; };
template <> struct RawOf<bf16> { typedef __half type; };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.01f * x; }

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace b200
