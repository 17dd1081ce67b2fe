// Host-side op wrappers: each UNETR layer type as one (or a few) launches.  T = activation type.
// CUDA-core engine (contract.cuh) here; the tcgen05 engine (tc_gemm.cuh) overrides the bf16 cases in exec.cu.
#pragma once
#include "contract.cuh"
#include "elementwise.cuh"

namespace b200 {

template <class T> struct Cl {  // channels-last window of a [N, D, H, W, pitch] tensor
  T* p; int pitch, coff, C;
};
template <class T> static inline Cl<T> cl(T* p, int pitch, int coff, int C) { Cl<T> c; c.p = p; c.pitch = pitch; c.coff = coff; c.C = C; return c; }
struct Sp { int N, D, H, W; long vox() const { return (long)D * H * W; } long rows() const { return (long)N * D * H * W; } };

template <class T, bool OF> static inline LdStrided<T, OF> ld2(const T* p, long so, long sk) {
  LdStrided<T, OF> l; l.p = p; l.so = so; l.sk = sk; l.sb0 = 0; l.sb1 = 0; l.nb1 = 1; return l;
}
template <class T, bool OF> static inline LdStrided<T, OF> ld4(const T* p, long so, long sk, long sb0, long sb1, int nb1) {
  LdStrided<T, OF> l; l.p = p; l.so = so; l.sk = sk; l.sb0 = sb0; l.sb1 = sb1; l.nb1 = nb1; return l;
}

// ---- linear layers (W fp32 [N,K], PyTorch layout) ----
template <class TA, class TO>
static int simt_linear_fwd(const TA* A, long lda, const float* W, int M, int N, int K, const EpStore<TO>& ep, cudaStream_t st) {
  B200_PROF("linear_fwd", st);
  return launch_contract(ld2<TA, false>(A, lda, 1), ld2<float, false>(W, K, 1), ep, M, N, K, 1, 1, st);
}
template <class TG, class TO>  // dX[M,K] = dY[M,N] W[N,K]
static int simt_linear_dgrad(const TG* dY, long ldy, const float* W, int M, int N, int K, const EpStore<TO>& ep, cudaStream_t st) {
  B200_PROF("linear_dgrad", st);
  return launch_contract(ld2<TG, false>(dY, ldy, 1), ld2<float, true>(W, 1, K), ep, M, K, N, 1, 1, st);
}
template <class TG, class TX>  // dW[N,K] = dY^T X
static int simt_linear_wgrad(const TG* dY, long ldy, const TX* X, long ldx, int M, int N, int K, float* dW, cudaStream_t st) {
  B200_PROF("linear_wgrad", st);
  return launch_contract(ld2<TG, true>(dY, 1, ldy), ld2<TX, true>(X, 1, ldx), ep_plain<float>(dW, K), N, K, M, 1, 1, st);
}

// ---- k^3 convolution, stride 1, same padding, bias-free (W fp32 [Co][Ci][k^3]) ----
template <class T>   // the output is a RAW conv output: stored as RawOf<T> (fp16 in bf16 mode)
static int simt_conv_fwd(Cl<const T> x, Sp sp, const float* W, int Co, int ks, Cl<T> out, cudaStream_t st) {
  B200_PROFD(st, "simt conv_fwd k%d %d->%d @%d", ks, x.C, Co, sp.D);
  typedef typename RawOf<T>::type TR;
  int taps = ks * ks * ks;
  RowIsOuter<ConvGather<T, true>, false> al; al.g = {x.p, sp.D, sp.H, sp.W, x.pitch, x.coff, x.C, ks, 1};
  ConvWeightB bl = {W, x.C, Co, taps, 0};
  EpStore<TR> ep = ep_plain<TR>(reinterpret_cast<TR*>(out.p) + out.coff, out.pitch);
  return launch_contract(al, bl, ep, (int)sp.rows(), Co, taps * x.C, 1, 1, st);
}
template <class T>
static int simt_conv_dgrad(Cl<const T> dy, Sp sp, const float* W, int Ci, int ks, Cl<T> dx, int accumulate, cudaStream_t st) {
  B200_PROFD(st, "simt conv_dgrad k%d %d->%d @%d", ks, dy.C, Ci, sp.D);
  int taps = ks * ks * ks;
  RowIsOuter<ConvGather<T, true>, false> al; al.g = {dy.p, sp.D, sp.H, sp.W, dy.pitch, dy.coff, dy.C, ks, -1};
  ConvWeightB bl = {W, Ci, dy.C, taps, 1};
  EpStore<T> ep = ep_plain<T>(dx.p + dx.coff, dx.pitch); ep.accumulate = accumulate;
  return launch_contract(al, bl, ep, (int)sp.rows(), Ci, taps * dy.C, 1, 1, st);
}
struct EpAtomicT {  // out[n*ld + m] += acc
  float* out; long ld;
  __device__ __forceinline__ void operator()(int, int m, int n, float acc) const { atomicAdd(out + (long)n * ld + m, acc); }
};
static inline int pick_splits(long K, long tiles) {
  long want = (148L * 4 + tiles - 1) / tiles;
  long maxs = (K + 1023) / 1024;
  long s = want < maxs ? want : maxs;
  return (int)(s < 1 ? 1 : s);
}
template <class T>  // dW[co][ci][tap] = sum_v dy[v,co] x[v+tap,ci]   (dW must be zero on entry)
static int simt_conv_wgrad(Cl<const T> x, Cl<const T> dy, Sp sp, int ks, float* dW, cudaStream_t st) {
  B200_PROFD(st, "simt conv_wgrad k%d %dx%d @%d", ks, x.C, dy.C, sp.D);
  int taps = ks * ks * ks;
  int Mr = x.C * taps, Nr = dy.C;
  RowIsK<ConvGather<T, false>, false> al; al.g = {x.p, sp.D, sp.H, sp.W, x.pitch, x.coff, x.C, ks, 1};
  LdStrided<T, true> bl = ld2<T, true>(dy.p + dy.coff, 1, dy.pitch);
  EpAtomicT ep = {dW, (long)Mr};
  long tiles = (long)cdiv(Mr, Nr <= 16 ? 256 : (Nr <= 32 ? 128 : 64)) * cdiv(Nr, Nr <= 16 ? 16 : (Nr <= 32 ? 32 : 64));
  return launch_contract(al, bl, ep, Mr, Nr, (int)sp.rows(), 1, pick_splits(sp.rows(), tiles), st);
}

// ---- ConvTranspose3d k2 s2 bias-free (W fp32 [Ci][Co][8]); sp = INPUT spatial dims ----
template <class TA, class T>
static int simt_convT_fwd(const TA* x, long ldx, int Ci, Sp sp, const float* W, Cl<T> out, cudaStream_t st) {
  B200_PROF("convT_fwd", st);
  EpConvTScatter<T> ep = {out.p, sp.D, sp.H, sp.W, out.pitch, out.coff};
  return launch_contract(ld2<TA, false>(x, ldx, 1), ld2<float, true>(W, 1, (long)out.C * 8), ep, (int)sp.rows(), out.C * 8, Ci, 1, 1, st);
}
template <class T, class TO>
static int simt_convT_dgrad(Cl<const T> dy, Sp sp, const float* W, int Ci, TO* dx, long lddx, int accumulate, cudaStream_t st) {
  B200_PROFD(st, "simt convT_dgrad %d<-%d @%d", Ci, dy.C, sp.D);
  RowIsOuter<ConvTGather<T>, false> al; al.g = {dy.p, sp.D, sp.H, sp.W, dy.pitch, dy.coff};
  EpStore<TO> ep = ep_plain<TO>(dx, lddx); ep.accumulate = accumulate;
  return launch_contract(al, ld2<float, false>(W, (long)dy.C * 8, 1), ep, (int)sp.rows(), Ci, dy.C * 8, 1, 1, st);
}
template <class TA, class T>  // dW[ci][co*8+tap] (dW zero on entry)
static int simt_convT_wgrad(const TA* x, long ldx, int Ci, Cl<const T> dy, Sp sp, float* dW, cudaStream_t st) {
  B200_PROFD(st, "simt convT_wgrad %dx%d @%d", Ci, dy.C, sp.D);
  RowIsK<ConvTGather<T>, true> bl; bl.g = {dy.p, sp.D, sp.H, sp.W, dy.pitch, dy.coff};
  int Nr = dy.C * 8;
  EpAtomic ep = {dW, (long)Nr};
  long tiles = (long)cdiv(Ci, 64) * cdiv(Nr, 64);
  return launch_contract(ld2<TA, true>(x, 1, ldx), bl, ep, Ci, Nr, (int)sp.rows(), 1, pick_splits(sp.rows(), tiles), st);
}

}  // namespace b200
