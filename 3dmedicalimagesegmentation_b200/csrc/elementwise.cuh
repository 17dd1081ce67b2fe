// HBM-bound kernels of the UNETR path: layout casts, LayerNorm, softmax, InstanceNorm(+LeakyReLU,+residual),
// column sums.  All take channels-last activations of type T (float = parity mode, bf16 = throughput mode),
// use 16-byte accesses and keep statistics in fp32/fp64.
#pragma once
#include "common.cuh"

namespace b200 {

struct ClView { long pitch; int coff; };  // element strides of a channels-last window

// ------------------------------------------------------------------ layout / casts
// dir 0: NCDHW fp32 -> channels-last T (dst[(n*V+v)*pitch+coff+c]); accumulate adds into dst.
// dir 1: channels-last T -> NCDHW fp32.
template <class T>
__global__ void layout_kernel(const float* __restrict__ ncdhw_in, float* __restrict__ ncdhw_out, T* cl, int C, long V,
                              int pitch, int coff, int dir, int accumulate, long total) {
  // 32x32 smem transpose over (c, v) tiles of one sample
  __shared__ float tile[32][33];
  long tiles_v = (V + 31) / 32; int tiles_c = (C + 31) / 32;
  long t = blockIdx.x; int n = (int)(t / (tiles_v * tiles_c)); long r = t % (tiles_v * tiles_c);
  int tc = (int)(r / tiles_v); long tv = r % tiles_v;
  int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (dir == 0) {
    for (int i = ty; i < 32; i += 8) {
      int c = tc * 32 + i; long v = tv * 32 + tx;
      tile[i][tx] = (c < C && v < V) ? ncdhw_in[((long)n * C + c) * V + v] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      long v = tv * 32 + i; int c = tc * 32 + tx;
      if (c < C && v < V) {
        long o = ((long)n * V + v) * pitch + coff + c;
        float val = tile[tx][i];
        if (accumulate) val += to_f(cl[o]);
        cl[o] = from_f<T>(val);
      }
    }
  } else {
    for (int i = ty; i < 32; i += 8) {
      long v = tv * 32 + i; int c = tc * 32 + tx;
      tile[i][tx] = (c < C && v < V) ? to_f(cl[((long)n * V + v) * pitch + coff + c]) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
      int c = tc * 32 + i; long v = tv * 32 + tx;
      if (c < C && v < V) ncdhw_out[((long)n * C + c) * V + v] = tile[tx][i];
    }
  }
}
template <class T>
static int launch_layout(const float* in, float* out, T* cl, int N, int C, long V, int pitch, int coff, int dir,
                         int accumulate, cudaStream_t st) {
  B200_PROF("layout", st);
  long blocks = (long)N * ((V + 31) / 32) * ((C + 31) / 32);
  layout_kernel<T><<<(unsigned)blocks, dim3(32, 8), 0, st>>>(in, out, cl, C, V, pitch, coff, dir, accumulate, 0);
  B200_LAUNCH_CHECK();
  return 0;
}

template <class TI, class TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long n) {
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = from_f<TO>(to_f(in[i]));
}
template <class TI, class TO>
static int launch_cast(const TI* in, TO* out, long n, cudaStream_t st) {
  B200_PROF("cast", st);
  int blocks = (int)min((long)148 * 8, (n + 255) / 256);
  B200_CUDA(launch_pdl(cast_kernel<TI, TO>, dim3(blocks), dim3(256), 0, st, in, out, n));
  B200_LAUNCH_CHECK();
  return 0;
}

struct CastJob { const float* src; bf16* dst; long n; };
struct CastJobs { CastJob j[64]; int count; };
// fp32 -> bf16 for up to 64 tensors in one launch: grid (blocks per tensor, tensor)
static __global__ void multi_cast_kernel(const CastJobs jobs) {
  const CastJob jb = jobs.j[blockIdx.y];
  long n4 = jb.n >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(jb.src);
  uint2* d2 = reinterpret_cast<uint2*>(jb.dst);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 v = s4[i];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    d2[i] = o;
  }
  for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < jb.n; i += (long)gridDim.x * blockDim.x)
    jb.dst[i] = __float2bfloat16_rn(jb.src[i]);
}

// Re-layout jobs of the bf16 engine in ONE launch (was one launch per conv layer):
//  kind 0: transposed conv  fp32 [Ci][Co][8] -> bf16 [Ci][8][Co]
//  kind 1: conv             fp32 [Co][Ci][taps] -> fwd bf16 [tap][co][ci] and dgrad bf16 [taps-1-tap][ci][co]
struct PackJob { const float* src; bf16* d0; bf16* d1; int a, b, taps, kind; };
struct PackJobs { PackJob j[32]; int count; };
// Both kinds go through a shared-memory tile so that global reads AND writes are runs of consecutive elements (the first form moved one
// element per thread with four 64-bit divisions and two 2-byte stores scattered at a stride of Co*Ci: 44 us for 7 M parameters).
//  kind 1: a block takes 16 co x 16 ci x taps: reads 16 runs of 16*taps floats, writes 32-byte runs over ci (fwd) and over co (dgrad)
//  kind 0: a block takes one ci row of Co*8 floats and writes it tap-major
static constexpr int kPackTile = 16 * 16 * 27;
static __global__ void __launch_bounds__(256) multi_pack_kernel(const PackJobs jobs) {
  const PackJob jb = jobs.j[blockIdx.y];
  __shared__ bf16 sm[kPackTile];
  if (jb.kind == 0) {
    const int Ci = jb.a, Co = jb.b, row = Co * 8;          // host checks row <= kPackTile
    for (int ci = blockIdx.x; ci < Ci; ci += gridDim.x) {
      const float* src = jb.src + (long)ci * row;
      for (int e = threadIdx.x; e < row; e += blockDim.x) sm[e] = __float2bfloat16_rn(src[e]);
      __syncthreads();
      bf16* dst = jb.d0 + (long)ci * row;
      for (int e = threadIdx.x; e < row; e += blockDim.x) { const int tap = e / Co, co = e - tap * Co; dst[e] = sm[co * 8 + tap]; }
      __syncthreads();
    }
  } else {
    const int Co = jb.a, Ci = jb.b, taps = jb.taps;         // host checks Co % 16 == 0, Ci % 16 == 0, taps <= 27
    const int tci = Ci >> 4, ntile = (Co >> 4) * tci, run = 16 * taps;
    for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
      const int co0 = (t / tci) << 4, ci0 = (t % tci) << 4;
      for (int e = threadIdx.x; e < 16 * run; e += blockDim.x) {
        const int col = e / run, j = e - col * run;          // j = ci_local * taps + tap: consecutive in the source
        sm[e] = __float2bfloat16_rn(jb.src[((long)(co0 + col) * Ci + ci0) * taps + j]);
      }
      __syncthreads();
      for (int e = threadIdx.x; e < 256 * taps; e += blockDim.x) {
        const int lo = e & 15, mid = (e >> 4) & 15, tap = e >> 8;
        // forward layout [tap][co][ci]: lo = ci, mid = co;  dgrad layout [taps-1-tap][ci][co]: lo = co, mid = ci
        jb.d0[((long)tap * Co + co0 + mid) * Ci + ci0 + lo] = sm[mid * run + lo * taps + tap];
        jb.d1[((long)(taps - 1 - tap) * Ci + ci0 + mid) * Co + co0 + lo] = sm[lo * run + mid * taps + tap];
      }
      __syncthreads();
    }
  }
}

static __global__ void add_kernel(float* __restrict__ dst, const float* __restrict__ src, long n) {
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] += src[i];
}
// dst += src; cast = T(dst)   (ViT backward: the encoder's gradient joins the residual stream at hidden states 3 / 6 / 9)
template <class T>
static __global__ void add_cast_kernel(float* __restrict__ dst, const float* __restrict__ src, T* __restrict__ cast, long n) {
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { const float v = dst[i] + src[i]; dst[i] = v; cast[i] = from_f<T>(v); }
}
template <class T>
static int launch_add_cast(float* dst, const float* src, T* cast, long n, cudaStream_t st) {
  B200_PROF("add", st);
  int blocks = (int)min((long)148 * 8, (n + 255) / 256);
  B200_CUDA(launch_pdl(add_cast_kernel<T>, dim3(blocks), dim3(256), 0, st, dst, src, cast, n));
  B200_LAUNCH_CHECK();
  return 0;
}
static int launch_add(float* dst, const float* src, long n, cudaStream_t st) {
  B200_PROF("add", st);
  int blocks = (int)min((long)148 * 8, (n + 255) / 256);
  B200_CUDA(launch_pdl(add_kernel, dim3(blocks), dim3(256), 0, st, dst, src, n));
  B200_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ LayerNorm (eps 1e-5, affine), one warp per row

// 4 consecutive elements of T <-> floats (16-byte / 8-byte accesses)
__device__ __forceinline__ void load4(const float* p, float* o) { float4 t = *reinterpret_cast<const float4*>(p); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }
__device__ __forceinline__ void load4(const bf16* p, float* o) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void store4(bf16* p, const float* v) {
  uint2 t; __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// Register-resident variants for H = NV4*128 (ViT-B: 768 -> NV4 = 6): every load of a row is issued before the first
// use, one pass over memory (the generic kernels below are latency-bound: 2-3 dependent passes of 24 scalar loads).
// Split-K producer fused into its consumer: x[row] = resid[row] + bias + sum_s part[s][row] is formed here (and written to xsum,
// the fp32 residual stream) instead of by atomics in the GEMM epilogue.
struct SplitSum { const float* part; int nsplit; long stride; const float* bias; const float* resid; float* xsum;
                  void* xsum_cast; };   // xsum_cast (nullable): a second copy of the summed row in the kernel's output type (the hidden states the encoders read)
template <class TO, int NV4>
__global__ void layernorm_fwd_reg_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                         TO* __restrict__ y, float* __restrict__ stats, int M, const SplitSum ss) {
  pdl_wait();
  constexpr int H = NV4 * 128;
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (long)row * H;
  float v[NV4][4];
  if (ss.nsplit) {
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const long o = (long)row * H + (i * 32 + lane) * 4;
      float b4[4], p4[4][4];   // <= 4 splits; fixed trip count + predication keeps all loads of the row independent (issued up front)
      load4(ss.resid + o, v[i]); load4(ss.bias + (i * 32 + lane) * 4, b4);
#pragma unroll
      for (int sp = 0; sp < 4; ++sp) {
        if (sp < ss.nsplit) load4(ss.part + sp * ss.stride + o, p4[sp]);
        else { p4[sp][0] = p4[sp][1] = p4[sp][2] = p4[sp][3] = 0.f; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] += b4[j] + ((p4[0][j] + p4[1][j]) + (p4[2][j] + p4[3][j]));
      store4(ss.xsum + o, v[i]);
      if (ss.xsum_cast) store4(reinterpret_cast<TO*>(ss.xsum_cast) + o, v[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV4; ++i) load4(xr + (i * 32 + lane) * 4, v[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
  float mean = warp_sum(s) * (1.f / H);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { float d = v[i][j] - mean; q = fmaf(d, d, q); }
  float rstd = rsqrtf(warp_sum(q) * (1.f / H) + 1e-5f);
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    float g[4], b[4], o[4];
    load4(gamma + (i * 32 + lane) * 4, g); load4(beta + (i * 32 + lane) * 4, b);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
    store4(y + (long)row * H + (i * 32 + lane) * 4, o);
  }
  if (lane == 0 && stats) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}
// ss.nsplit > 0: the incoming gradient g is the sum of the producer GEMM's split-K partials (fp32); it is also written back as TG to
// ss_gout (the parameter-gradient kernel reads it)
template <class TG, int NV4>
__global__ void layernorm_bwd_dx_reg_kernel(const TG* __restrict__ g, const float* __restrict__ x, const float* __restrict__ stats,
                                            const float* __restrict__ gamma, const float* dx_res, float* dx_out, TG* dx_out_cast, int M,
                                            const SplitSum ss, TG* ss_gout) {
  pdl_wait();
  constexpr int H = NV4 * 128;
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  float gg[NV4][4], xh[NV4][4], rs[NV4][4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const long o = (long)row * H + (i * 32 + lane) * 4;
    float gm[4];
    if (ss.nsplit) {
      float p4[3][4];
      load4(ss.part + o, gg[i]);
#pragma unroll
      for (int sp = 1; sp < 4; ++sp) {
        if (sp < ss.nsplit) load4(ss.part + sp * ss.stride + o, p4[sp - 1]);
        else { p4[sp - 1][0] = p4[sp - 1][1] = p4[sp - 1][2] = p4[sp - 1][3] = 0.f; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) gg[i][j] += p4[0][j] + (p4[1][j] + p4[2][j]);
      store4(ss_gout + o, gg[i]);
    } else {
      load4(g + o, gg[i]);
    }
    load4(x + o, xh[i]); load4(gamma + (i * 32 + lane) * 4, gm);
    if (dx_res) load4(dx_res + o, rs[i]);
    else { rs[i][0] = rs[i][1] = rs[i][2] = rs[i][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { gg[i][j] *= gm[j]; xh[i][j] = (xh[i][j] - mean) * rstd; }
  }
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { s1 += gg[i][j]; s2 = fmaf(gg[i][j], xh[i][j], s2); }
  s1 = warp_sum(s1) * (1.f / H); s2 = warp_sum(s2) * (1.f / H);
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const long o = (long)row * H + (i * 32 + lane) * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = rstd * (gg[i][j] - s1 - xh[i][j] * s2) + rs[i][j];
    store4(dx_out + o, v);
    if (dx_out_cast) store4(dx_out_cast + o, v);
  }
}
template <class TO>
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, TO* __restrict__ y, float* __restrict__ stats,
                                     int M, int H) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (long)row * H;
  float s = 0.f;
  for (int i = lane; i < H; i += 32) s += xr[i];
  float mean = warp_sum(s) / H;
  float q = 0.f;
  for (int i = lane; i < H; i += 32) { float d = xr[i] - mean; q += d * d; }
  float rstd = rsqrtf(warp_sum(q) / H + 1e-5f);
  for (int i = lane; i < H; i += 32) y[(long)row * H + i] = from_f<TO>((xr[i] - mean) * rstd * gamma[i] + beta[i]);
  if (lane == 0 && stats) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}
template <class TO>
static int launch_layernorm_fwd(const float* x, const float* g, const float* b, TO* y, float* stats, int M, int H,
                                cudaStream_t st, const SplitSum* ss = nullptr) {
  B200_PROF("layernorm_fwd", st);
  SplitSum none; memset(&none, 0, sizeof(none));
  B200_CHECK(!ss || H == 768, "fused split-K LayerNorm needs hidden size 768");
  if (H == 768) B200_CUDA(launch_pdl(layernorm_fwd_reg_kernel<TO, 6>, dim3(cdiv(M, 4)), dim3(128), 0, st, x, g, b, y, stats, M, ss ? *ss : none));
  else layernorm_fwd_kernel<TO><<<cdiv(M, 8), 256, 0, st>>>(x, g, b, y, stats, M, H);
  B200_LAUNCH_CHECK();
  return 0;
}

// dx_out = dx_res (nullable) + rstd*(g*gamma - mean(g*gamma) - xhat*mean(g*gamma*xhat))
template <class TG>
__global__ void layernorm_bwd_dx_kernel(const TG* __restrict__ g, const float* __restrict__ x,
                                        const float* __restrict__ stats, const float* __restrict__ gamma,
                                        const float* dx_res, float* dx_out, TG* dx_out_cast, int M, int H) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= M) return;
  float mean = stats[2 * row], rstd = stats[2 * row + 1];
  const float* xr = x + (long)row * H;
  const TG* gr = g + (long)row * H;
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < H; i += 32) {
    float gg = to_f(gr[i]) * gamma[i]; float xh = (xr[i] - mean) * rstd;
    s1 += gg; s2 += gg * xh;
  }
  s1 = warp_sum(s1) / H; s2 = warp_sum(s2) / H;
  for (int i = lane; i < H; i += 32) {
    float gg = to_f(gr[i]) * gamma[i]; float xh = (xr[i] - mean) * rstd;
    float v = rstd * (gg - s1 - xh * s2);
    if (dx_res) v += dx_res[(long)row * H + i];
    dx_out[(long)row * H + i] = v;
    if (dx_out_cast) dx_out_cast[(long)row * H + i] = from_f<TG>(v);
  }
}
// dgamma[h] = sum_rows g*xhat ; dbeta[h] = sum_rows g.   grid = H/32 blocks of (32 x 8)
template <class TG>
__global__ void layernorm_bwd_params_kernel(const TG* __restrict__ g, const float* __restrict__ x,
                                            const float* __restrict__ stats, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta, int M, int H) {
  pdl_wait();
  __shared__ float sg[32][33], sb[32][33];
  int col = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, b = 0.f;
  if (col < H) {
#pragma unroll 4
    for (int r = threadIdx.y; r < M; r += 32) {
      float gg = to_f(g[(long)r * H + col]);
      a += gg * (x[(long)r * H + col] - stats[2 * r]) * stats[2 * r + 1];
      b += gg;
    }
  }
  sg[threadIdx.y][threadIdx.x] = a; sb[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && col < H) {
#pragma unroll
    for (int i = 1; i < 32; ++i) { a += sg[i][threadIdx.x]; b += sb[i][threadIdx.x]; }
    if (dgamma) dgamma[col] = a;
    if (dbeta) dbeta[col] = b;
  }
}
template <class TG>
static int launch_layernorm_bwd(const TG* g, const float* x, const float* stats, const float* gamma,
                                const float* dx_res, float* dx_out, TG* dx_out_cast, float* dgamma, float* dbeta, int M, int H,
                                cudaStream_t st, const SplitSum* ss = nullptr) {
  B200_PROF("layernorm_bwd", st);
  SplitSum none; memset(&none, 0, sizeof(none));
  B200_CHECK(!ss || H == 768, "fused split-K LayerNorm needs hidden size 768");
  if (H == 768) B200_CUDA(launch_pdl(layernorm_bwd_dx_reg_kernel<TG, 6>, dim3(cdiv(M, 4)), dim3(128), 0, st, g, x, stats, gamma, dx_res, dx_out, dx_out_cast, M, ss ? *ss : none, const_cast<TG*>(g)));
  else layernorm_bwd_dx_kernel<TG><<<cdiv(M, 8), 256, 0, st>>>(g, x, stats, gamma, dx_res, dx_out, dx_out_cast, M, H);
  B200_LAUNCH_CHECK();
  if (dgamma || dbeta) {
    B200_CUDA(launch_pdl(layernorm_bwd_params_kernel<TG>, dim3(cdiv(H, 32)), dim3(dim3(32, 32)), 0, st, g, x, stats, dgamma, dbeta, M, H));
    B200_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------ column sums (bias gradients)
template <class TG>
__global__ void colsum_kernel(const TG* __restrict__ g, float* __restrict__ out, int M, int N) {
  pdl_wait();
  __shared__ float s[32][33];
  int col = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (col < N) {
#pragma unroll 4
    for (int r = threadIdx.y; r < M; r += 32) a += to_f(g[(long)r * N + col]);
  }
  s[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
#pragma unroll
    for (int i = 1; i < 32; ++i) a += s[i][threadIdx.x];
    out[col] = a;
  }
}
template <class TG>
static int launch_colsum(const TG* g, float* out, int M, int N, cudaStream_t st) {
  B200_PROF("colsum", st);
  B200_CUDA(launch_pdl(colsum_kernel<TG>, dim3(cdiv(N, 32)), dim3(dim3(32, 32)), 0, st, g, out, M, N));
  B200_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ batched parameter-gradient reductions
// The ViT backward defers everything that only feeds the optimizer (weight, bias and LayerNorm-parameter gradients) and issues it per
// group of transformer blocks: ONE launch for all bias column sums of the group and ONE for all LayerNorm parameter sums, instead of
// 3 + 2 launches of 24..96 blocks per transformer block (each ~4-5 us of pure latency on a 1.3 MB tensor).
static constexpr int kMaxRedJobs = 40;
struct ColsumJob { const void* g; float* out; int M, N; };
struct ColsumJobs { ColsumJob j[kMaxRedJobs]; int count; };
template <class TG>
__global__ void colsum_multi_kernel(const ColsumJobs jobs) {
  pdl_wait();
  const ColsumJob jb = jobs.j[blockIdx.y];
  const int col = blockIdx.x * 32 + threadIdx.x;
  if (blockIdx.x * 32 >= jb.N) return;
  __shared__ float s[32][33];
  const TG* g = reinterpret_cast<const TG*>(jb.g);
  float a = 0.f;
  if (col < jb.N) {
#pragma unroll 4
    for (int r = threadIdx.y; r < jb.M; r += 32) a += to_f(g[(long)r * jb.N + col]);
  }
  s[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && col < jb.N) {
#pragma unroll
    for (int i = 1; i < 32; ++i) a += s[i][threadIdx.x];
    jb.out[col] = a;
  }
}
template <class TG>
static int launch_colsum_multi(const ColsumJobs& jobs, cudaStream_t st) {
  if (jobs.count <= 0) return 0;
  B200_PROF("colsum", st);
  int maxN = 0; for (int i = 0; i < jobs.count; ++i) maxN = jobs.j[i].N > maxN ? jobs.j[i].N : maxN;
  B200_CUDA(launch_pdl(colsum_multi_kernel<TG>, dim3(cdiv(maxN, 32), jobs.count), dim3(32, 32), 0, st, jobs));
  B200_LAUNCH_CHECK();
  return 0;
}

struct LnParamJob { const void* g; const float* x; const float* stats; float* dgamma; float* dbeta; };
struct LnParamJobs { LnParamJob j[kMaxRedJobs]; int count, M, H; };
template <class TG>
__global__ void layernorm_bwd_params_multi_kernel(const LnParamJobs jobs) {
  pdl_wait();
  __shared__ float sg[32][33], sb[32][33];
  const LnParamJob jb = jobs.j[blockIdx.y];
  const int M = jobs.M, H = jobs.H;
  const TG* g = reinterpret_cast<const TG*>(jb.g);
  const int col = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, b = 0.f;
  if (col < H) {
#pragma unroll 4
    for (int r = threadIdx.y; r < M; r += 32) {
      const float gg = to_f(g[(long)r * H + col]);
      a += gg * (jb.x[(long)r * H + col] - jb.stats[2 * r]) * jb.stats[2 * r + 1];
      b += gg;
    }
  }
  sg[threadIdx.y][threadIdx.x] = a; sb[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && col < H) {
#pragma unroll
    for (int i = 1; i < 32; ++i) { a += sg[i][threadIdx.x]; b += sb[i][threadIdx.x]; }
    if (jb.dgamma) jb.dgamma[col] = a;
    if (jb.dbeta) jb.dbeta[col] = b;
  }
}
template <class TG>
static int launch_layernorm_bwd_params_multi(const LnParamJobs& jobs, cudaStream_t st) {
  if (jobs.count <= 0) return 0;
  B200_PROF("layernorm_bwd_params", st);
  B200_CUDA(launch_pdl(layernorm_bwd_params_multi_kernel<TG>, dim3(cdiv(jobs.H, 32), jobs.count), dim3(32, 32), 0, st, jobs));
  B200_LAUNCH_CHECK();
  return 0;
}

// out[row % C] += sum_v x[row, v]   (x is [rows, V] fp32; bias gradient of the NCDHW head)
static __global__ void rowsum_atomic_kernel(const float* __restrict__ x, float* __restrict__ out, long V, int C) {
  int row = blockIdx.y;
  long per = (V + gridDim.x - 1) / gridDim.x;
  long v0 = (long)blockIdx.x * per, v1 = min(V, v0 + per);
  float a = 0.f;
  for (long v = v0 + threadIdx.x; v < v1; v += blockDim.x) a += x[(long)row * V + v];
  a = warp_sum(a);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s[i];
    atomicAdd(out + row % C, t);
  }
}

// Pixel-unshuffle of the 2x up-sampled channels-last tensor y (window coff..coff+Co of pitch) into the GEMM operand
// U[v_in][tap*Co + co]  (tap = (a*2+b)*2+c of the k2s2 transposed conv).  16-byte moves both ways.
// IDX = unsigned when the element count fits 32 bits (every shape of the benchmark): eight div/mod per 16-byte move are then 32-bit.
template <class T, class IDX>
__global__ void unshuffle_kernel(const T* __restrict__ y, ClView yv, int Co, int N, int D, int H, int W, T* __restrict__ U) {
  pdl_wait();
  constexpr int VN = Vec16<T>::N;
  const IDX lanes = (IDX)(Co / VN);
  const IDX total = (IDX)N * D * H * W * 8 * lanes;
  for (IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (IDX)gridDim.x * blockDim.x) {
    const IDX r = e / lanes; const int lv = (int)(e - r * lanes); const int tap = (int)(r & 7); const IDX v = r >> 3;
    IDX t = v / (IDX)W; const int w = (int)(v - t * (IDX)W);
    IDX t2 = t / (IDX)H; const int h = (int)(t - t2 * (IDX)H);
    const IDX n = t2 / (IDX)D; const int d = (int)(t2 - n * (IDX)D);
    long pos = (((long)n * 2 * D + 2 * d + (tap >> 2)) * 2 * H + 2 * h + ((tap >> 1) & 1)) * 2 * W + 2 * w + (tap & 1);
    Vec16<T> a; a.load(y + pos * yv.pitch + yv.coff + lv * VN);
    a.store(U + (long)v * 8 * Co + tap * Co + lv * VN);
  }
}
// fp32 [Ci][Co][8] -> bf16 [Ci][8][Co]  (tap-major copy of a transposed-conv weight)
static __global__ void pack_convT_tapmajor_kernel(const float* __restrict__ W, bf16* __restrict__ out, int Ci, int Co) {
  long total = (long)Ci * Co * 8;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int tap = (int)(e & 7); long r = e >> 3; int co = (int)(r % Co); long ci = r / Co;
    out[(ci * 8 + tap) * Co + co] = __float2bfloat16_rn(W[e]);
  }
}
// bf16 patch rows A[tok][k] from the NCDHW fp32 volume (same k order as the weight: perceptron or conv)
static __global__ void patch_gather_kernel(const float* __restrict__ x, bf16* __restrict__ A, int C, int S0, int S1, int S2, int g0, int g1, int g2,
                                    int conv_order, long total) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    int Kp = 4096 * C; int k = (int)(e % Kp); int tok = (int)(e / Kp);
    int c, p;
    if (conv_order) { c = k >> 12; p = k & 4095; } else { c = k % C; p = k / C; }
    int p3 = p & 15, p2 = (p >> 4) & 15, p1 = p >> 8;
    int d = tok % g2; int t = tok / g2; int w = t % g1; t /= g1; int h = t % g0; int b = t / g0;
    A[e] = __float2bfloat16_rn(x[((((long)b * C + c) * S0 + h * 16 + p1) * S1 + w * 16 + p2) * S2 + d * 16 + p3]);
  }
}

// The same rows when k runs over (c, p1, p2, p3) with p3 fastest (pos_embed == "conv", or one input channel, where the two orders
// coincide): a thread converts 8 consecutive voxels of a patch line (two 16-byte loads, one 16-byte store) with 32-bit index arithmetic.
static __global__ void __launch_bounds__(256) patch_gather8_kernel(const float* __restrict__ x, bf16* __restrict__ A, int C, int S0, int S1, int S2, int g0, int g1, int g2,
                                                                   int total8) {
  const int Kp8 = 512 * C;                       // groups of 8 per token row
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total8; e += gridDim.x * blockDim.x) {
    const int tok = e / Kp8, k = (e - tok * Kp8) << 3;
    const int c = k >> 12, p = k & 4095, p3 = p & 15, p2 = (p >> 4) & 15, p1 = p >> 8;
    const int d = tok % g2; int t = tok / g2; const int w = t % g1; t /= g1; const int h = t % g0, b = t / g0;
    const float* src = x + ((((long)b * C + c) * S0 + h * 16 + p1) * S1 + w * 16 + p2) * S2 + d * 16 + p3;
    const float4 u = *reinterpret_cast<const float4*>(src), v = *reinterpret_cast<const float4*>(src + 4);
    __nv_bfloat162 q0 = __floats2bfloat162_rn(u.x, u.y), q1 = __floats2bfloat162_rn(u.z, u.w), q2 = __floats2bfloat162_rn(v.x, v.y), q3 = __floats2bfloat162_rn(v.z, v.w);
    uint4 o; o.x = *reinterpret_cast<uint32_t*>(&q0); o.y = *reinterpret_cast<uint32_t*>(&q1); o.z = *reinterpret_cast<uint32_t*>(&q2); o.w = *reinterpret_cast<uint32_t*>(&q3);
    *reinterpret_cast<uint4*>(A + ((long)e << 3)) = o;
  }
}

// dpos[l,h] = sum_b dx[b,l,h]
static __global__ void batchsum_kernel(const float* __restrict__ dx, float* __restrict__ out, int B, long LH) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= LH) return;
  float a = 0.f;
  for (int b = 0; b < B; ++b) a += dx[(long)b * LH + i];
  out[i] = a;
}

// ------------------------------------------------------------------ attention softmax (scale applied before softmax)
// S fp32 [rows, ld] (first L columns valid) -> P (type TP) [rows, ld]; one warp per row.
template <class TP>
__global__ void softmax_fwd_kernel(const float* __restrict__ S, TP* __restrict__ P, long rows, int L, int ld, float scale) {
  pdl_wait();
  long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* s = S + row * ld;
  float mx = -INFINITY;
  for (int i = lane; i < L; i += 32) mx = fmaxf(mx, s[i] * scale);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int i = lane; i < L; i += 32) sum += __expf(s[i] * scale - mx);
  sum = warp_sum(sum);
  float inv = 1.f / sum;
  for (int i = lane; i < ld; i += 32) P[row * ld + i] = from_f<TP>(i < L ? __expf(s[i] * scale - mx) * inv : 0.f);
}
// dS = P * (dP - sum_j dP_j P_j) * scale
template <class TP>
__global__ void softmax_bwd_kernel(const TP* __restrict__ P, const float* __restrict__ dP, TP* __restrict__ dS, long rows,
                                   int L, int ld, float scale) {
  pdl_wait();
  long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float dot = 0.f;
  for (int i = lane; i < L; i += 32) dot += to_f(P[row * ld + i]) * dP[row * ld + i];
  dot = warp_sum(dot);
  for (int i = lane; i < ld; i += 32)
    dS[row * ld + i] = from_f<TP>(i < L ? to_f(P[row * ld + i]) * (dP[row * ld + i] - dot) * scale : 0.f);
}

// ------------------------------------------------------------------ InstanceNorm3d(affine=False, eps 1e-5, biased var)
// Tensors are channels-last [N, V, pitch] with channel window [coff, coff+C).  C % Vec16<T>::N == 0,
// C/VecN a power of two <= 256.  Statistics: double sum/sumsq -> float (mean, rstd) per (n,c).



// Block reduction of NVAL per-thread floats over all threads that share (threadIdx.x % lanes): xor-shuffles inside the warp,
// then one smem row per (warp, lane<lanes); returns through `red` ([warps][lanes][NVAL]) after a __syncthreads().
// (The previous tail -- `lanes` threads serially adding 128 x NVAL smem values in double -- cost ~45 us per block.)
template <int NVAL>
__device__ __forceinline__ void block_reduce_groups(float* vals, int lanes, float* red) {
  for (int o = lanes; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < NVAL; ++i) vals[i] += __shfl_xor_sync(0xffffffffu, vals[i], o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < lanes) {
#pragma unroll
    for (int i = 0; i < NVAL; ++i) red[((warp * lanes) + lane) * NVAL + i] = vals[i];
  }
  __syncthreads();
}

template <class T>
__global__ void in_stats_kernel(const typename RawOf<T>::type* __restrict__ x, ClView xv, int C, long V, double* __restrict__ acc) {
  pdl_wait();
  typedef typename RawOf<T>::type TR;
  constexpr int VN = Vec16<T>::N;
  extern __shared__ float red[];  // [256][2*VN]
  int lanes = C / VN;             // vectors per voxel
  int lv = threadIdx.x % lanes, sub = threadIdx.x / lanes, nsub = blockDim.x / lanes;
  int n = blockIdx.y;
  long per = (V + gridDim.x - 1) / gridDim.x;
  long v0 = (long)blockIdx.x * per, v1 = min(V, v0 + per);
  float s[VN], q[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) s[i] = q[i] = 0.f;
#pragma unroll 4
  for (long v = v0 + sub; v < v1; v += nsub) {
    Vec16<TR> a; a.load(x + ((long)n * V + v) * xv.pitch + xv.coff + lv * VN);
#pragma unroll
    for (int i = 0; i < VN; ++i) { s[i] += a.v[i]; q[i] += a.v[i] * a.v[i]; }
  }
  float vals[2 * VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) { vals[i] = s[i]; vals[VN + i] = q[i]; }
  const int lanes_eff = lanes < 32 ? lanes : 32;   // lanes > 32: a warp covers only part of the channels
  if (lanes <= 32) {
    block_reduce_groups<2 * VN>(vals, lanes, red);
    const int nwarp = blockDim.x >> 5;
    for (int e = threadIdx.x; e < lanes_eff * 2 * VN; e += blockDim.x) {
      int l = e / (2 * VN), i = e % (2 * VN);
      double tot = 0.0;
      for (int wv = 0; wv < nwarp; ++wv) tot += red[(wv * lanes + l) * 2 * VN + i];
      atomicAdd(acc + ((long)n * C + l * VN + i % VN) * 2 + i / VN, tot);
    }
  } else {
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      atomicAdd(acc + ((long)n * C + lv * VN + i) * 2, (double)s[i]);
      atomicAdd(acc + ((long)n * C + lv * VN + i) * 2 + 1, (double)q[i]);
    }
  }
}
// (mean, rstd) of one (n, c) from the double (sum, sumsq) pair that the conv epilogues / in_stats_kernel accumulate.  There is no separate
// finalize launch: every block of the consuming normalise pass derives the constants of its own channels from the sums, and block 0 of
// each sample also stores them as (mean, rstd) for the backward (17 launches of one CTA each per forward otherwise).
__device__ __forceinline__ void in_moments(const double* __restrict__ acc2, double invV, float& mean, float& rstd) {
  double m = acc2[0] * invV, var = acc2[1] * invV - m * m;
  if (var < 0) var = 0;
  mean = (float)m; rstd = 1.0f / sqrtf((float)var + 1e-5f);
}

// out = lrelu(norm(x))                            (two==0)
// out = lrelu(norm_a(x) + norm_b(x2))             (two==1)
// acc1 / acc2 (nullable): the (sum, sumsq) accumulators of x / x2 -- the constants come from them (in_moments) and are stored to mr / mr2
// by block 0; null: mr / mr2 already hold (mean, rstd)
template <class T>
__global__ void in_apply_kernel(const typename RawOf<T>::type* __restrict__ x, ClView xv, float* __restrict__ mr,
                                const typename RawOf<T>::type* __restrict__ x2, ClView x2v, float* __restrict__ mr2, T* __restrict__ out,
                                ClView ov, int C, long V, int two, const double* __restrict__ acc1, const double* __restrict__ acc2, double invV) {
  pdl_wait();
  constexpr int VN = Vec16<T>::N;
  int lanes = C / VN;
  int n = blockIdx.y;
  long total = V * lanes;
  // lanes is a power of two dividing the grid stride, so this thread always owns the same VN channels: constants live in registers
  const long e0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lgl = 31 - __clz(lanes);   // lanes is a power of two
  const int lv = (int)(e0 & (lanes - 1)), c0 = lv * VN;
  // one channel per thread: (mean, rstd) from the sums (or from memory) into shared memory; block 0 keeps them for the backward
  __shared__ __align__(16) float smr[4][256];   // mean1, rstd1, mean2, rstd2 per channel (C <= 256)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const long ch = (long)n * C + c;
    float m, r;
    if (acc1) { in_moments(acc1 + 2 * ch, invV, m, r); if (blockIdx.x == 0) { mr[2 * ch] = m; mr[2 * ch + 1] = r; } }
    else { m = mr[2 * ch]; r = mr[2 * ch + 1]; }
    smr[0][c] = m; smr[1][c] = r;
    m = 0.f; r = 0.f;
    if (two) {
      if (acc2) { in_moments(acc2 + 2 * ch, invV, m, r); if (blockIdx.x == 0) { mr2[2 * ch] = m; mr2[2 * ch + 1] = r; } }
      else { m = mr2[2 * ch]; r = mr2[2 * ch + 1]; }
    }
    smr[2][c] = m; smr[3][c] = r;
  }
  __syncthreads();
  float m1[VN], r1[VN], m2[VN], r2[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) { m1[i] = smr[0][c0 + i]; r1[i] = smr[1][c0 + i]; m2[i] = smr[2][c0 + i]; r2[i] = smr[3][c0 + i]; }
#pragma unroll 4
  for (long e = e0; e < total; e += (long)gridDim.x * blockDim.x) {
    long v = e >> lgl;
    Vec16<typename RawOf<T>::type> a; a.load(x + ((long)n * V + v) * xv.pitch + xv.coff + c0);
    Vec16<T> o;
    if (two) {
      Vec16<typename RawOf<T>::type> b; b.load(x2 + ((long)n * V + v) * x2v.pitch + x2v.coff + c0);
#pragma unroll
      for (int i = 0; i < VN; ++i) o.v[i] = lrelu((a.v[i] - m1[i]) * r1[i] + (b.v[i] - m2[i]) * r2[i]);
    } else {
#pragma unroll
      for (int i = 0; i < VN; ++i) o.v[i] = lrelu((a.v[i] - m1[i]) * r1[i]);
    }
    o.store(out + ((long)n * V + v) * ov.pitch + ov.coff + c0);
  }
}

// Backward, pass 1: per-(n,c) sums.  g = dOut * lrelu'(act).
//  TWO == false: act = lrelu(n1) is the saved activation; n1 recovered from it.   sums: [Sg, Sg*n1]
//  TWO == true : act = lrelu(n2+n3); RAW moments [Sg, S g*c2, S g*c3] of the raw conv outputs are accumulated (no per-channel
//                constants in the loop) and converted to [Sg, Sg*n2, Sg*n3] in the prologue of in_bwd_apply_kernel: Sg*n = rstd*(S g*c - mean*Sg)
template <class T, bool TWO, bool RESIGN = false>
__global__ void __launch_bounds__(256, 3) in_bwd_reduce_kernel(const T* __restrict__ dout, ClView dv, const T* __restrict__ act, ClView av,
                                     const typename RawOf<T>::type* __restrict__ ra, ClView rav, const typename RawOf<T>::type* __restrict__ rb, ClView rbv, int C, long V,
                                     double* __restrict__ acc /*[N][C][3]*/, const float* __restrict__ mra, const float* __restrict__ mrb) {
  pdl_wait();
  constexpr int VN = Vec16<T>::N;
  extern __shared__ float red[];  // [warps][lanes][3*VN]
  int lanes = C / VN;
  int lv = threadIdx.x % lanes, sub = threadIdx.x / lanes, nsub = blockDim.x / lanes;
  int n = blockIdx.y, c0 = lv * VN;
  long per = (V + gridDim.x - 1) / gridDim.x;
  long v0 = (long)blockIdx.x * per, v1 = min(V, v0 + per);
  float s0[VN], s1[VN], s2[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) s0[i] = s1[i] = s2[i] = 0.f;
  // RESIGN (TWO only; act is not read): the sign of the block output n2 + n3 = c2*r2 + c3*r3 - (m2*r2 + m3*r3) is recomputed from the raw conv outputs
  // instead of reading the stored activation (one tensor pass less); the three constants per channel sit in shared memory
  constexpr bool resign = TWO && RESIGN;
  __shared__ __align__(16) float kc[TWO ? 3 * 256 : 4];
  if (resign) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const long ch = (long)n * C + c;
      const float qa = mra[2 * ch + 1], qb = mrb[2 * ch + 1];
      kc[c] = qa; kc[C + c] = qb; kc[2 * C + c] = -(mra[2 * ch] * qa + mrb[2 * ch] * qb);
    }
    __syncthreads();
  }
#pragma unroll 4
  for (long v = v0 + sub; v < v1; v += nsub) {
    long base = (long)n * V + v;
    Vec16<T> d, a; d.load(dout + base * dv.pitch + dv.coff + c0);
    if (!resign) a.load(act + base * av.pitch + av.coff + c0);
    if (TWO) {
      Vec16<typename RawOf<T>::type> xa, xb; xa.load(ra + base * rav.pitch + rav.coff + c0); xb.load(rb + base * rbv.pitch + rbv.coff + c0);
#pragma unroll
      for (int i = 0; i < VN; i += 4) {
        float qa[4] = {0.f, 0.f, 0.f, 0.f}, qb[4] = {0.f, 0.f, 0.f, 0.f}, kk[4] = {0.f, 0.f, 0.f, 0.f};
        if (resign) {
          const float4 A = *reinterpret_cast<const float4*>(kc + c0 + i), B = *reinterpret_cast<const float4*>(kc + C + c0 + i), K = *reinterpret_cast<const float4*>(kc + 2 * C + c0 + i);
          qa[0] = A.x; qa[1] = A.y; qa[2] = A.z; qa[3] = A.w; qb[0] = B.x; qb[1] = B.y; qb[2] = B.z; qb[3] = B.w; kk[0] = K.x; kk[1] = K.y; kk[2] = K.z; kk[3] = K.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float pre = resign ? fmaf(xa.v[i + q], qa[q], fmaf(xb.v[i + q], qb[q], kk[q])) : a.v[i + q];
          float g = d.v[i + q] * (pre > 0.f ? 1.f : 0.01f);
          s0[i + q] += g; s1[i + q] = fmaf(g, xa.v[i + q], s1[i + q]); s2[i + q] = fmaf(g, xb.v[i + q], s2[i + q]);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        float g = d.v[i] * (a.v[i] > 0.f ? 1.f : 0.01f);
        float nn = a.v[i] > 0.f ? a.v[i] : a.v[i] * 100.f;
        s0[i] += g; s1[i] = fmaf(g, nn, s1[i]);
      }
    }
  }
  float vals[3 * VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) { vals[i] = s0[i]; vals[VN + i] = s1[i]; vals[2 * VN + i] = s2[i]; }
  if (lanes <= 32) {
    block_reduce_groups<3 * VN>(vals, lanes, red);
    const int nwarp = blockDim.x >> 5;
    for (int e = threadIdx.x; e < lanes * 3 * VN; e += blockDim.x) {
      int l = e / (3 * VN), i = e % (3 * VN);
      if (!TWO && i >= 2 * VN) continue;
      double tot = 0.0;
      for (int wv = 0; wv < nwarp; ++wv) tot += red[(wv * lanes + l) * 3 * VN + i];
      atomicAdd(acc + ((long)n * C + l * VN + i % VN) * 3 + i / VN, tot);
    }
  } else {
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      double* a3 = acc + ((long)n * C + c0 + i) * 3;
      atomicAdd(a3, (double)s0[i]); atomicAdd(a3 + 1, (double)s1[i]);
      if (TWO) atomicAdd(a3 + 2, (double)s2[i]);
    }
  }
}
// Backward, pass 2:  d(raw) = rstd * (g - mean(g) - n * mean(g*n))  =  A1*g + A2*raw + A3  with per-(n,c) constants staged in smem
//   A1 = rstd, A2 = -rstd^2 * mean(g n), A3 = rstd^2 * mean(g n) * mean - rstd * mean(g)      (TWO: one triple per raw input)
//   !TWO: n is recovered from the saved activation: d = rstd * (g - mg - n*mgn)
template <class T, bool TWO, bool RESIGN = false>
__global__ void __launch_bounds__(256, 3) in_bwd_apply_kernel(const T* __restrict__ dout, ClView dv, const T* __restrict__ act, ClView av,
                                    const typename RawOf<T>::type* __restrict__ ra, ClView rav, const float* __restrict__ mra,
                                    const typename RawOf<T>::type* __restrict__ rb, ClView rbv, const float* __restrict__ mrb, int C, long V,
                                    const double* __restrict__ acc, T* __restrict__ da, ClView dav,
                                    T* __restrict__ db, ClView dbv) {
  pdl_wait();
  constexpr int VN = Vec16<T>::N;
  extern __shared__ __align__(16) float cst[];   // [7][C]: 6 affine constants + the sign constant (RESIGN, see in_bwd_reduce_kernel)
  int lanes = C / VN;
  int n = blockIdx.y;
  long total = V * lanes;
  float invV = 1.f / (float)V;
  constexpr bool resign = TWO && RESIGN;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double* a3 = acc + ((long)n * C + c) * 3;
    float mg = (float)a3[0] * invV, mga = (float)a3[1] * invV;
    float m = mra[((long)n * C + c) * 2], r = mra[((long)n * C + c) * 2 + 1];
    if (TWO) {   // acc holds RAW moments here (in_bwd_reduce_kernel<TWO>): centre / normalise them first, in double like the sums
      float m2 = mrb[((long)n * C + c) * 2], r2 = mrb[((long)n * C + c) * 2 + 1];
      mga = (float)((double)r * (a3[1] - (double)m * a3[0])) * invV;
      float mgb = (float)((double)r2 * (a3[2] - (double)m2 * a3[0])) * invV;
      cst[c] = r; cst[C + c] = -r * r * mga; cst[2 * C + c] = r * r * mga * m - r * mg;
      cst[3 * C + c] = r2; cst[4 * C + c] = -r2 * r2 * mgb; cst[5 * C + c] = r2 * r2 * mgb * m2 - r2 * mg;
      cst[6 * C + c] = -(m * r + m2 * r2);
    } else {
      cst[c] = r; cst[C + c] = -r * mga; cst[2 * C + c] = -r * mg;   // d = r*g + (-r*mga)*n + (-r*mg)
    }
  }
  __syncthreads();
  const long e0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lgl = 31 - __clz(lanes);   // lanes is a power of two
  const int lv = (int)(e0 & (lanes - 1)), c0 = lv * VN;   // fixed per thread (lanes divides the grid stride)
#pragma unroll 2
  for (long e = e0; e < total; e += (long)gridDim.x * blockDim.x) {
    long v = e >> lgl;
    long base = (long)n * V + v;
    Vec16<T> d, a; d.load(dout + base * dv.pitch + dv.coff + c0);
    if (!resign) a.load(act + base * av.pitch + av.coff + c0);
    Vec16<T> oa, ob;
    if (TWO) {
      Vec16<typename RawOf<T>::type> xa, xb; xa.load(ra + base * rav.pitch + rav.coff + c0); xb.load(rb + base * rbv.pitch + rbv.coff + c0);
#pragma unroll
      for (int i = 0; i < VN; i += 4) {
        float4 A1 = *reinterpret_cast<const float4*>(cst + c0 + i), A2 = *reinterpret_cast<const float4*>(cst + C + c0 + i), A3 = *reinterpret_cast<const float4*>(cst + 2 * C + c0 + i);
        float4 B1 = *reinterpret_cast<const float4*>(cst + 3 * C + c0 + i), B2 = *reinterpret_cast<const float4*>(cst + 4 * C + c0 + i), B3 = *reinterpret_cast<const float4*>(cst + 5 * C + c0 + i);
        const float a1[4] = {A1.x, A1.y, A1.z, A1.w}, a2[4] = {A2.x, A2.y, A2.z, A2.w}, a3c[4] = {A3.x, A3.y, A3.z, A3.w};
        const float b1[4] = {B1.x, B1.y, B1.z, B1.w}, b2[4] = {B2.x, B2.y, B2.z, B2.w}, b3[4] = {B3.x, B3.y, B3.z, B3.w};
        float kk[4] = {0.f, 0.f, 0.f, 0.f};   // -(m2*r2 + m3*r3): n2 + n3 = c2*r2 + c3*r3 + kk (resign only; r2 = a1, r3 = b1)
        if (resign) { const float4 K = *reinterpret_cast<const float4*>(cst + 6 * C + c0 + i); kk[0] = K.x; kk[1] = K.y; kk[2] = K.z; kk[3] = K.w; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float pre = resign ? fmaf(xa.v[i + q], a1[q], fmaf(xb.v[i + q], b1[q], kk[q])) : a.v[i + q];
          float g = d.v[i + q] * (pre > 0.f ? 1.f : 0.01f);
          oa.v[i + q] = fmaf(a1[q], g, fmaf(a2[q], xa.v[i + q], a3c[q]));
          ob.v[i + q] = fmaf(b1[q], g, fmaf(b2[q], xb.v[i + q], b3[q]));
        }
      }
      oa.store(da + base * dav.pitch + dav.coff + c0);
      ob.store(db + base * dbv.pitch + dbv.coff + c0);
    } else {
#pragma unroll
      for (int i = 0; i < VN; i += 4) {
        float4 A1 = *reinterpret_cast<const float4*>(cst + c0 + i), A2 = *reinterpret_cast<const float4*>(cst + C + c0 + i), A3 = *reinterpret_cast<const float4*>(cst + 2 * C + c0 + i);
        const float a1[4] = {A1.x, A1.y, A1.z, A1.w}, a2[4] = {A2.x, A2.y, A2.z, A2.w}, a3c[4] = {A3.x, A3.y, A3.z, A3.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float g = d.v[i + q] * (a.v[i + q] > 0.f ? 1.f : 0.01f);
          float na = a.v[i + q] > 0.f ? a.v[i + q] : a.v[i + q] * 100.f;
          oa.v[i + q] = fmaf(a1[q], g, fmaf(a2[q], na, a3c[q]));
        }
      }
      oa.store(da + base * dav.pitch + dav.coff + c0);
    }
  }
}

// ~4 16-byte vectors per thread, at most 8 blocks per SM; `lanes` = vectors per voxel (C / VecN).  (V/512 left the 24^3 and 12^3
// levels with 27-54 blocks: 15-25 us of pure latency for a few MB.)
static inline int in_grid_x(long V, int lanes = 2) { return (int)max(1L, min((long)148 * 4, V * lanes / 1024)); }

}  // namespace b200
