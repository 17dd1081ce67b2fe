// UNETR forward / backward executor.  Mirrors unetr.py:182-208 (wiring) and the MONAI 0.6.0 block semantics
// recorded in SURVEY.md Appendix B, but on a B200-first data layout:
//   * tokens [B*L, hidden] row-major ARE the channels-last 16x-downsampled volume, so `proj_feat`
//     (unetr.py:177-180) costs nothing;
//   * every decoder tensor is channels-last (NDHWC) so a voxel's channels are one contiguous GEMM row;
//   * `torch.cat((up, skip), 1)` is never materialised: the transposed conv writes channels [0,C) and the skip
//     branch writes [C,2C) of one concat buffer;
//   * the residual stream and all statistics are fp32, activations are T (float = parity mode, bf16 = throughput).
// Caller owns all memory: parameters / gradients are PyTorch-layout fp32 pointers, the workspace is one buffer.
#pragma once
#include <type_traits>
#include <functional>
#include <vector>

#include "ops.cuh"
#include "edge_kernels.cuh"
#include "tc_gemm.cuh"
#include "tc_gemm_grouped.cuh"
#include "tc_attention.cuh"
#include "tc_conv.cuh"
#include "tc_conv_halo.cuh"
#include "tc_wgrad_halo.cuh"
#include "tc_wgrad.cuh"

#include "exec_iface.h"

namespace b200 {


enum ParamIdx {
  P_POS = 0, P_PATCH_W, P_PATCH_B, P_BLK0 = 3,
  // per transformer block (11): LN1_W LN1_B QKV_W PROJ_W PROJ_B LN2_W LN2_B FC1_W FC1_B FC2_W FC2_B
  P_NORM_W = 3 + 12 * 11, P_NORM_B,
  P_E1_C1, P_E1_C2, P_E1_C3,
  P_E2_T0, P_E2_T1, P_E2_T2,
  P_E3_T0, P_E3_T1,
  P_E4_T0,
  P_D5_T, P_D5_C1, P_D5_C2, P_D5_C3,
  P_D4_T, P_D4_C1, P_D4_C2, P_D4_C3,
  P_D3_T, P_D3_C1, P_D3_C2, P_D3_C3,
  P_D2_T, P_D2_C1, P_D2_C2, P_D2_C3,
  P_OUT_W, P_OUT_B, P_COUNT
};
enum { B_LN1_W = 0, B_LN1_B, B_QKV_W, B_PROJ_W, B_PROJ_B, B_LN2_W, B_LN2_B, B_FC1_W, B_FC1_B, B_FC2_W, B_FC2_B, B_COUNT };

enum { FLAG_NEED_ENCODER_GRAD = 1, FLAG_HAS_DLOGITS = 2, FLAG_HAS_DENC4 = 4, FLAG_SAVE_FOR_BACKWARD = 8, FLAG_NO_BACKWARD = 16,   // forward only: no backward call follows (inference)
       FLAG_INPLACE_WGRADS = 32,   // backward with gradient events: keep the conv-stack weight gradients in place (otherwise: deferred behind the ViT backward)
       FLAG_WEIGHTS_PACKED = 64 };   // the bf16 weight copies in this workspace are current (same workspace, unchanged parameters)

struct Bump {
  char* base; size_t off;
  template <class U> U* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    U* p = base ? reinterpret_cast<U*>(base + off) : nullptr;
    off += n * sizeof(U);
    return p;
  }
};

template <class T>
struct ResSave {  // what a residual conv block keeps for its backward
  T *a1, *c2, *c3; float *mr1, *mr2, *mr3;
};

template <class T>
struct Workspace {
  // ViT
  float *x0, *hs[12], *x1[12], *ln1s[12], *ln2s[12], *lnfs, *S;
  T *ln1[12], *qkv[12], *P[12], *att[12], *ln2[12], *u[12], *h[12], *vit_out, *hsT[3];
  // decoder
  T *xcl, *c1tmp, *e2a, *e2b, *e3a, *cat5, *cat4, *cat3, *cat2, *d3, *d2, *d1, *d0;
  ResSave<T> rs[5];  // enc1, dec5, dec4, dec3, dec2
  double* stat_acc;      // current (sum, sumsq) accumulator pair: a fresh pre-zeroed slot of stat_pool per use (one memset per forward)
  double* stat_pool; int stat_next;
  // bf16 mode only: packed bf16 copies of the GEMM weights (refreshed every forward), same layouts as the fp32 masters
  bf16 *wqkv[12], *wproj[12], *wfc1[12], *wfc2[12], *wT[10];
  bf16 *wcf[15], *wcd[15];
  bf16 *wTt[10], *wpatch, *apatch;  // tap-major transposed-conv weights; patch weight; gathered bf16 patch rows [M, 4096*Cin]  // conv weights packed [tap][co][ci] (forward) and [tap'][ci][co] (dgrad, taps flipped)
  // backward scratch
  float *dx, *dx2, *dhs[3], *dP;
  float* part;   // split-K partial tiles [<=4][M][H] (summed by the LayerNorm kernel that consumes the GEMM)
  float* wg_scratch; size_t wg_scratch_bytes;   // partial tiles of the deterministic weight-gradient epilogue (tc_wgrad.cuh)
  T *unsh;  // unsh: pixel-unshuffled dOut of a transposed conv [rows_in, 8*Co]  // bf16 mode: operand copies of the fp32 residual-stream gradients
  T *dvit, *datt, *dS, *gA, *dcat, *dc2, *dc3, *da1, *dc1;
  T* dcs[5][3];          // per residual block (indexed like rs): dc1, dc2, dc3 kept until the deferred weight-gradient launches (Exec::defer_wg)
  // per-block operands of the DEFERRED parameter gradients (weight / bias / LayerNorm-parameter sums are issued per group of blocks,
  // see vit_param_grads): dyo[i+1] = d(hs[i]) and dyo[0] = d(x0) as T, dh[i] = d(fc1 pre-activation), dy1[i] = d(x1[i]) as T,
  // dqkv[i], dln2[i] / dln1[i] = gradients wrt the two LayerNorm outputs
  T *dyo[13], *dh[12], *dy1[12], *dqkv[12], *dln2[12], *dln1[12];
  double* bwd_acc;       // current backward-norm accumulator: a fresh pre-zeroed slot of bwd_pool per use (one memset per backward)
  double* bwd_pool; int bwd_next;
  size_t bytes;
};

template <class T>
struct Exec {
  typedef typename RawOf<T>::type TR;   // storage type of raw conv outputs (c1, c2, c3 of every residual block)
  static constexpr int kStatSlots = 20, kBwdSlots = 12;   // >= accumulator uses per forward / backward (5 blocks x 3 / x 2)
  int next_stat() { if (w.stat_next >= kStatSlots) { set_error("statistics slot pool exhausted"); return 1; } w.stat_acc = w.stat_pool + (size_t)(w.stat_next++) * 4 * c.B * 8 * c.fs; return 0; }
  int next_bwd() { if (w.bwd_next >= kBwdSlots) { set_error("backward accumulator pool exhausted"); return 1; } w.bwd_acc = w.bwd_pool + (size_t)(w.bwd_next++) * 3 * c.B * 8 * c.fs; return 0; }
  UnetrConfig c;
  int g0, g1, g2, L, Lp, M, H, F, nh, dh;
  long V[5];  // voxels per sample at levels 0 (full) .. 4 (tokens)
  Workspace<T> w;
  static constexpr bool kTC = std::is_same<T, bf16>::value;  // tcgen05 engine for the dense contractions

  // the 10 transposed convs in table order: (param index, Ci, Co)
  void convT_desc(int i, int& pidx, int& Ci, int& Co) const {
    const int fs = c.fs, Hh = c.hidden;
    const int tab[10][3] = {{P_E2_T0, Hh, 2 * fs}, {P_E2_T1, 2 * fs, 2 * fs}, {P_E2_T2, 2 * fs, 2 * fs}, {P_E3_T0, Hh, 4 * fs},
                            {P_E3_T1, 4 * fs, 4 * fs}, {P_E4_T0, Hh, 8 * fs}, {P_D5_T, Hh, 8 * fs}, {P_D4_T, 8 * fs, 4 * fs},
                            {P_D3_T, 4 * fs, 2 * fs}, {P_D2_T, 2 * fs, fs}};
    pidx = tab[i][0]; Ci = tab[i][1]; Co = tab[i][2];
  }
  // the 15 convolutions in table order: (param index, Ci, Co, kernel)
  void conv_desc(int i, int& pidx, int& Ci, int& Co, int& ks) const {
    const int fs = c.fs;
    const int base[5] = {P_E1_C1, P_D5_C1, P_D4_C1, P_D3_C1, P_D2_C1};
    const int cin[5] = {c.Cin, 16 * fs, 8 * fs, 4 * fs, 2 * fs}, cout[5] = {fs, 8 * fs, 4 * fs, 2 * fs, fs};
    int blk = i / 3, which = i % 3;
    pidx = base[blk] + which; Co = cout[blk]; Ci = which == 1 ? cout[blk] : cin[blk]; ks = which == 2 ? 1 : 3;
  }
  int conv_index(const float* W) const {
    if (!cur_params) return -1;
    for (int i = 0; i < 15; ++i) { int p, ci, co, ks; conv_desc(i, p, ci, co, ks); if (cur_params[p] == W) return i; }
    return -1;
  }
  // element count of the conv-side parameters (indices P_E1_C1 .. P_OUT_B), 0 for others
  size_t conv_param_elems(int pidx) const {
    for (int i = 0; i < 15; ++i) { int p, ci, co, ks; conv_desc(i, p, ci, co, ks); if (p == pidx) return (size_t)ci * co * ks * ks * ks; }
    for (int i = 0; i < 10; ++i) { int p, ci, co; convT_desc(i, p, ci, co); if (p == pidx) return (size_t)ci * co * 8; }
    if (pidx == P_OUT_W) return (size_t)c.ncls * c.fs;
    if (pidx == P_OUT_B) return (size_t)c.ncls;
    return 0;
  }
  // All atomically accumulated weight gradients (convs, transposed convs, head) are zeroed by ONE memset when they are
  // consecutive slices of one buffer (the Python host hands out views of a flat gradient buffer); otherwise per tensor.
  bool grads_prezeroed = false;
  int prezero_conv_grads(float* const* G, cudaStream_t st) {
    grads_prezeroed = false;
    float* lo = nullptr; float* expect = nullptr; size_t total = 0;
    for (int i = P_E1_C1; i < P_COUNT; ++i) {
      if (!G[i]) continue;
      size_t n = conv_param_elems(i);
      if (!lo) { lo = G[i]; expect = lo; }
      if (G[i] != expect) return 0;          // not contiguous: the per-tensor memsets stay in charge
      expect = G[i] + n; total += n;
    }
    if (!lo) return 0;
    B200_CUDA(cudaMemsetAsync(lo, 0, sizeof(float) * total, st));
    grads_prezeroed = true;
    return 0;
  }
  int zero_grad(float* p, size_t n, cudaStream_t st) {
    if (!grads_prezeroed) B200_CUDA(cudaMemsetAsync(p, 0, sizeof(float) * n, st));
    return 0;
  }
  size_t convT_elems(int i) const { int p, ci, co; convT_desc(i, p, ci, co); return (size_t)ci * co * 8; }
  const bf16* convT_packed(const float* const* P, const float* W, bool tap_major = false) const {
    for (int i = 0; i < 10; ++i) { int p, ci, co; convT_desc(i, p, ci, co); if (P[p] == W) return tap_major ? w.wTt[i] : w.wT[i]; }
    return nullptr;
  }

  explicit Exec(const UnetrConfig& cfg) : c(cfg) {
    g0 = c.S0 / 16; g1 = c.S1 / 16; g2 = c.S2 / 16;
    L = g0 * g1 * g2; Lp = (L + 7) & ~7; M = c.B * L; H = c.hidden; F = c.mlp; nh = c.heads; dh = H / nh;
    V[4] = L;
    for (int l = 3; l >= 0; --l) V[l] = V[l + 1] * 8;
  }
  Sp sp(int level) const { int s = 16 >> level; return Sp{c.B, g0 * s, g1 * s, g2 * s}; }

  void layout(char* base, bool with_backward) {
    Bump b{base, 0};
    size_t MH = (size_t)M * H, MF = (size_t)M * F, PP = (size_t)c.B * nh * L * Lp;
    int fs = c.fs, B = c.B;
    w.x0 = b.take<float>(MH);
    w.part = b.take<float>(4 * MH);
    for (int i = 0; i < 12; ++i) {
      w.hs[i] = b.take<float>(MH); w.x1[i] = b.take<float>(MH);
      w.ln1s[i] = b.take<float>(2 * M); w.ln2s[i] = b.take<float>(2 * M);
      w.ln1[i] = b.take<T>(MH); w.qkv[i] = b.take<T>(3 * MH); w.P[i] = b.take<T>(PP); w.att[i] = b.take<T>(MH);
      w.ln2[i] = b.take<T>(MH); w.u[i] = b.take<T>(MF); w.h[i] = b.take<T>(MF);
    }
    w.lnfs = b.take<float>(2 * M); w.S = b.take<float>(PP); w.vit_out = b.take<T>(MH);
    for (int i = 0; i < 3; ++i) w.hsT[i] = b.take<T>(MH);
    w.xcl = b.take<T>((size_t)B * V[0] * c.Cin);
    w.c1tmp = b.take<T>((size_t)B * V[0] * fs);
    w.e2a = b.take<T>((size_t)B * V[3] * 2 * fs); w.e2b = b.take<T>((size_t)B * V[2] * 2 * fs);
    w.e3a = b.take<T>((size_t)B * V[3] * 4 * fs);
    w.cat5 = b.take<T>((size_t)B * V[3] * 16 * fs); w.cat4 = b.take<T>((size_t)B * V[2] * 8 * fs);
    w.cat3 = b.take<T>((size_t)B * V[1] * 4 * fs); w.cat2 = b.take<T>((size_t)B * V[0] * 2 * fs);
    w.d3 = b.take<T>((size_t)B * V[3] * 8 * fs); w.d2 = b.take<T>((size_t)B * V[2] * 4 * fs);
    w.d1 = b.take<T>((size_t)B * V[1] * 2 * fs); w.d0 = b.take<T>((size_t)B * V[0] * fs);
    const int lvl[5] = {0, 3, 2, 1, 0}; const int co[5] = {fs, 8 * fs, 4 * fs, 2 * fs, fs};
    for (int i = 0; i < 5; ++i) {
      size_t n = (size_t)B * V[lvl[i]] * co[i];
      w.rs[i].a1 = b.take<T>(n); w.rs[i].c2 = b.take<T>(n); w.rs[i].c3 = b.take<T>(n);
      w.rs[i].mr1 = b.take<float>(2 * B * co[i]); w.rs[i].mr2 = b.take<float>(2 * B * co[i]); w.rs[i].mr3 = b.take<float>(2 * B * co[i]);
    }
    w.stat_pool = b.take<double>((size_t)kStatSlots * 4 * B * 8 * fs);   // slots of two (sum, sumsq) accumulators
    w.stat_acc = w.stat_pool; w.stat_next = 0;
    if (kTC) w.apatch = b.take<bf16>((size_t)M * 4096 * c.Cin);
    if (with_backward) {
      w.dx = b.take<float>(MH); w.dx2 = b.take<float>(MH);
      for (int i = 0; i < 3; ++i) w.dhs[i] = b.take<float>(MH);
      w.dP = b.take<float>(PP); w.dS = b.take<T>(PP);
      w.dvit = b.take<T>(MH); w.datt = b.take<T>(MH);
      for (int i = 0; i < 13; ++i) w.dyo[i] = b.take<T>(MH);
      for (int i = 0; i < 12; ++i) {
        w.dh[i] = b.take<T>(MF); w.dy1[i] = b.take<T>(MH); w.dqkv[i] = b.take<T>(3 * MH);
        w.dln2[i] = b.take<T>(MH); w.dln1[i] = b.take<T>(MH);
      }
      size_t big = (size_t)B * V[0] * fs;
      w.gA = b.take<T>(big); w.dcat = b.take<T>(2 * big); w.dc2 = b.take<T>(big); w.dc3 = b.take<T>(big);
      w.da1 = b.take<T>(big); w.dc1 = b.take<T>(big); w.unsh = b.take<T>(big);
      { const int lvl[5] = {0, 3, 2, 1, 0};
        for (int k = 0; k < 5; ++k) for (int j = 0; j < 3; ++j) w.dcs[k][j] = b.take<T>((size_t)B * V[lvl[k]] * (fs << lvl[k])); }
      w.wg_scratch_bytes = 0;
      if constexpr (kTC) {
        const int lvl_of_blk[5] = {0, 3, 2, 1, 0};
        for (int i = 0; i < 15; ++i) {
          int pp, ci, co, ks; conv_desc(i, pp, ci, co, ks);
          if (!tc::wgrad_supported(ci, co, 8, 0, 8, 0) || tc::wgrad_halo_supported(ci, co, ks)) continue;
          Sp s = sp(lvl_of_blk[i / 3]);
          size_t need = tc::wgrad_scratch_bytes(ci, co, ks, s.N, s.D, s.H, s.W);
          if (need > w.wg_scratch_bytes) w.wg_scratch_bytes = need;
        }
      }
      w.wg_scratch = b.take<float>(w.wg_scratch_bytes / sizeof(float));
      w.bwd_pool = b.take<double>((size_t)kBwdSlots * 3 * B * 8 * fs);
      w.bwd_acc = w.bwd_pool; w.bwd_next = 0;
    }
    w.bytes = (b.off + 255) & ~(size_t)255;
  }

  // bf16 mode: the packed bf16 copies of the GEMM / conv weights live in ONE caller-owned buffer that outlives the workspaces (it does
  // not depend on the batch size): the forward refreshes it (pack_weights) unless the caller says it is current
  // (FLAG_WEIGHTS_PACKED) -- FusedAdamW writes the copies itself while it updates the fp32 masters (adamw.cuh), so a training
  // step has no cast / re-layout launch at all.  Returns the byte size; base == nullptr only measures.
  char* packed_base = nullptr;
  size_t layout_packed(char* base) {
    Bump b{base, 0};
    if (kTC) {
      for (int i = 0; i < 12; ++i) {
        w.wqkv[i] = b.take<bf16>((size_t)3 * H * H); w.wproj[i] = b.take<bf16>((size_t)H * H);
        w.wfc1[i] = b.take<bf16>((size_t)F * H); w.wfc2[i] = b.take<bf16>((size_t)H * F);
      }
      for (int i = 0; i < 10; ++i) { w.wT[i] = b.take<bf16>(convT_elems(i)); w.wTt[i] = b.take<bf16>(convT_elems(i)); }
      w.wpatch = b.take<bf16>((size_t)H * 4096 * c.Cin);
      for (int i = 0; i < 15; ++i) { int p, ci, co, ks; conv_desc(i, p, ci, co, ks); size_t n = (size_t)ci * co * ks * ks * ks; w.wcf[i] = b.take<bf16>(n); w.wcd[i] = b.take<bf16>(n); }
    }
    return (b.off + 255) & ~(size_t)255;
  }
  // Byte offset of the PLAIN bf16 cast of parameter `pidx` in the packed buffer, or -1 when the parameter has none (biases, norms,
  // position embedding, and the conv weights, which only exist re-laid-out).  These are the copies FusedAdamW writes itself.
  long long packed_cast_offset(int pidx) {
    if (!kTC) return -1;
    char* const fake = reinterpret_cast<char*>(uintptr_t(1) << 20);      // offsets only: nothing is dereferenced; forward()/backward() re-bind
    layout_packed(fake);
    auto off = [fake](const void* q) { return (long long)(reinterpret_cast<const char*>(q) - fake); };
    if (pidx == P_PATCH_W) return off(w.wpatch);
    if (pidx >= P_BLK0 && pidx < P_NORM_W) {
      const int i = (pidx - P_BLK0) / B_COUNT, k = (pidx - P_BLK0) % B_COUNT;
      const bf16* q = k == B_QKV_W ? w.wqkv[i] : k == B_PROJ_W ? w.wproj[i] : k == B_FC1_W ? w.wfc1[i] : k == B_FC2_W ? w.wfc2[i] : nullptr;
      return q ? off(q) : -1;
    }
    for (int i = 0; i < 10; ++i) { int pp, ci, co; convT_desc(i, pp, ci, co); if (pp == pidx) return off(w.wT[i]); }
    return -1;
  }

  // ------------------------------------------------------------ engine dispatch (CUDA-core engine; see exec.cu for tcgen05)
  // Wb = packed bf16 copy of W (bf16 mode) -- operands of the tcgen05 engine; fp32 mode runs the CUDA-core engine.
  template <class TO>
  int linear_fwd(const T* A, long lda, const float* W, const bf16* Wb, int Mr, int N, int K, const EpStore<TO>& ep, cudaStream_t st) {
    if constexpr (kTC) {
      B200_PROFD(st, "linear_fwd %dx%dx%d", Mr, N, K);
      return tc::gemm(tc::operand(A, lda, 1), tc::operand(Wb, K, 1), ep, Mr, N, K, 1, 1, st);
    } else {
      return simt_linear_fwd(A, lda, W, Mr, N, K, ep, st);
    }
  }
  // Split-K variants for the N = hidden GEMMs (48 output tiles on 148 SMs): the splits store raw fp32 partial tiles into w.part and
  // the LayerNorm kernel that consumes the result sums them (+ bias + residual).  Returns the split count in *ss (0 = not split).
  bool can_split(int Mr, int N, int K) const { return kTC && H == 768 && N == H && Mr == M && tc::plan_splitk(Mr, N, K, 1) > 1; }
  int linear_fwd_split(const T* A, long lda, const bf16* Wb, int Mr, int N, int K, const float* bias, const float* resid, float* xsum,
                       SplitSum* ss, cudaStream_t st) {
    if constexpr (kTC) {
      B200_PROFD(st, "linear_fwd %dx%dx%d", Mr, N, K);
      const int ks = tc::plan_splitk(Mr, N, K, 1);
      EpStore<float> e = ep_plain<float>(w.part, N); e.splitk_nbat = 1; e.split_stride = (long)Mr * N;
      B200_TRY(tc::gemm(tc::operand(A, lda, 1), tc::operand(Wb, K, 1), e, Mr, N, K, 1, 1, st, false, ks));
      *ss = SplitSum{w.part, ks, (long)Mr * N, bias, resid, xsum, nullptr};
    }
    return 0;
  }
  int linear_dgrad_split(const T* dY, long ldy, const bf16* Wb, int Mr, int N, int K, SplitSum* ss, cudaStream_t st) {
    if constexpr (kTC) {
      B200_PROFD(st, "linear_dgrad %dx%dx%d", Mr, K, N);
      const int ks = tc::plan_splitk(Mr, K, N, 1);
      EpStore<float> e = ep_plain<float>(w.part, K); e.splitk_nbat = 1; e.split_stride = (long)Mr * K;
      B200_TRY(tc::gemm(tc::operand(dY, ldy, 1), tc::operand(Wb, 1, K), e, Mr, K, N, 1, 1, st, false, ks));
      *ss = SplitSum{w.part, ks, (long)Mr * K, nullptr, nullptr, nullptr};
    }
    return 0;
  }
  template <class TO>  // dX[M,K] = dY[M,N] W[N,K]   (W is the MN-major B operand: no transposed copy)
  int linear_dgrad(const T* dY, long ldy, const float* W, const bf16* Wb, int Mr, int N, int K, const EpStore<TO>& ep, cudaStream_t st) {
    if constexpr (kTC) {
      B200_PROFD(st, "linear_dgrad %dx%dx%d", Mr, K, N);
      return tc::gemm(tc::operand(dY, ldy, 1), tc::operand(Wb, 1, K), ep, Mr, K, N, 1, 1, st);
    } else {
      return simt_linear_dgrad(dY, ldy, W, Mr, N, K, ep, st);
    }
  }
  // dW[N,K] = dY^T X   (both operands MN-major)
  int linear_wgrad(const T* dY, long ldy, const T* X, long ldx, int Mr, int N, int K, float* dW, cudaStream_t st) {
    if constexpr (kTC) {
      B200_PROFD(st, "linear_wgrad %dx%dx%d", N, K, Mr);
      return tc::gemm(tc::operand(dY, 1, ldy), tc::operand(X, 1, ldx), ep_plain<float>(dW, K), N, K, Mr, 1, 1, st);
    } else {
      return simt_linear_wgrad(dY, ldy, X, ldx, Mr, N, K, dW, st);
    }
  }
  // bf16 mode: refresh the packed weight copies (one launch for all ViT + transposed-conv weights)
  // casts = false: only the re-laid-out conv / transposed-conv copies (the plain casts are kept current by FusedAdamW)
  int pack_weights(const float* const* P, cudaStream_t st, bool casts = true) {
    if constexpr (kTC) {
      B200_PROF("pack_weights", st);
      if (casts) {
      CastJobs jobs; int n = 0;
      auto add = [&](const float* src, bf16* dst, size_t cnt) { jobs.j[n].src = src; jobs.j[n].dst = dst; jobs.j[n].n = (long)cnt; ++n; };
      for (int i = 0; i < 12; ++i) {
        const float* const* bp = P + P_BLK0 + i * B_COUNT;
        add(bp[B_QKV_W], w.wqkv[i], (size_t)3 * H * H); add(bp[B_PROJ_W], w.wproj[i], (size_t)H * H);
        add(bp[B_FC1_W], w.wfc1[i], (size_t)F * H); add(bp[B_FC2_W], w.wfc2[i], (size_t)H * F);
      }
      for (int i = 0; i < 10; ++i) { int p, ci, co; convT_desc(i, p, ci, co); add(P[p], w.wT[i], convT_elems(i)); }
      add(P[P_PATCH_W], w.wpatch, (size_t)H * 4096 * c.Cin);
      jobs.count = n;
      multi_cast_kernel<<<dim3(64, n), 256, 0, st>>>(jobs);
      B200_LAUNCH_CHECK();
      }
      PackJobs pj; int m = 0;
      for (int i = 0; i < 10; ++i) {
        int p, ci, co; convT_desc(i, p, ci, co);
        pj.j[m++] = PackJob{P[p], w.wTt[i], nullptr, ci, co, 8, 0};
      }
      for (int i = 0; i < 15; ++i) {
        int p, ci, co, ks; conv_desc(i, p, ci, co, ks);
        if (ci % 16 || co % 16) continue;   // those layers stay on the CUDA-core engine
        pj.j[m++] = PackJob{P[p], w.wcf[i], w.wcd[i], co, ci, ks * ks * ks, 1};
      }
      pj.count = m;
      for (int i = 0; i < m; ++i)
        B200_CHECK(pj.j[i].kind == 0 ? pj.j[i].b * 8 <= kPackTile : pj.j[i].taps <= 27, "weight re-layout: tile does not fit (job %d)", i);
      // grid.x = tiles of the largest job (256 -> 128 channels: 8 x 16 tiles of 16 co x 16 ci x 27 taps); smaller jobs' surplus blocks exit at once
      multi_pack_kernel<<<dim3(128, m), 256, 0, st>>>(pj);
      B200_LAUNCH_CHECK();
    }
    return 0;
  }

  // ------------------------------------------------------------ InstanceNorm helpers
  // (sum, sumsq) accumulators whose (mean, rstd) have not been derived yet: the normalise pass that consumes `mr` does it in its
  // prologue (elementwise.cuh: in_moments) instead of a finalize launch per statistic
  struct PendStat { const float* mr; const double* acc; } pend_stat[4] = {};
  void set_pending(const float* mr, const double* acc) {
    for (auto& p : pend_stat) if (!p.mr || p.mr == mr) { p.mr = mr; p.acc = acc; return; }
    pend_stat[0].mr = mr; pend_stat[0].acc = acc;   // unreachable: at most three statistics are in flight per residual block
  }
  const double* take_pending(const float* mr) {
    if (mr) for (auto& p : pend_stat) if (p.mr == mr) { p.mr = nullptr; return p.acc; }
    return nullptr;
  }
  int in_stats(Cl<const T> x, long Vs, float* mr, cudaStream_t st, bool have_sums = false) {
    if (have_sums) { set_pending(mr, w.stat_acc); return 0; }   // sums already accumulated by the conv epilogue
    B200_PROF("instnorm_stats", st);
    constexpr int VN = Vec16<T>::N;
    B200_CHECK(x.C % VN == 0 && 256 % (x.C / VN) == 0 && x.pitch % VN == 0 && x.coff % VN == 0,
               "InstanceNorm channel count %d unsupported (need a power of two >= 8)", x.C);
    B200_TRY(next_stat());
    dim3 g(in_grid_x(Vs, x.C / VN), c.B);
    B200_CUDA(launch_pdl(in_stats_kernel<T>, dim3(g), dim3(256), 256 * 2 * VN * sizeof(float), st, reinterpret_cast<const TR*>(x.p), ClView{x.pitch, x.coff}, x.C, Vs, w.stat_acc));
    B200_LAUNCH_CHECK();
    set_pending(mr, w.stat_acc);
    return 0;
  }
  int in_apply(Cl<const T> x, float* mr, const T* x2, float* mr2, Cl<T> out, long Vs, cudaStream_t st) {
    B200_PROF("instnorm_apply", st);
    B200_CHECK(x.C <= 256, "InstanceNorm over %d channels unsupported (<= 256)", x.C);
    dim3 g(in_grid_x(Vs, x.C / Vec16<T>::N) * 2, c.B);
    const double* a1 = take_pending(mr); const double* a2 = take_pending(mr2);
    B200_CUDA(launch_pdl(in_apply_kernel<T>, dim3(g), dim3(256), 0, st, reinterpret_cast<const TR*>(x.p), ClView{x.pitch, x.coff}, mr, reinterpret_cast<const TR*>(x2), ClView{x.C, 0}, mr2, out.p,
                                          ClView{out.pitch, out.coff}, x.C, Vs, (int)(x2 != nullptr), a1, a2, 1.0 / (double)Vs));
    B200_LAUNCH_CHECK();
    return 0;
  }

  // ------------------------------------------------------------ residual conv block (UnetResBlock, Appendix B.2)
  // conv ops are virtual-ish hooks so the tcgen05 engine can replace them
  // stats != null asks for the InstanceNorm sums of the output; *stats_done tells whether the conv epilogue produced them
  int conv_fwd(Cl<const T> x, Sp s, const float* W, int Co, int ks, Cl<T> out, double* stats, bool* stats_done, cudaStream_t st) {
    if (stats_done) *stats_done = false;
    if constexpr (kTC) {
      int ci = conv_index(W);
      if (ci >= 0 && tc::conv_supported(x.C, Co, x.pitch, x.coff, out.pitch, out.coff)) {
        B200_PROFD(st, "conv_fwd k%d %d->%d @%d", ks, x.C, Co, s.D);
        if (stats) { B200_TRY(next_stat()); stats = w.stat_acc; if (stats_done) *stats_done = true; }
        if (ks == 3 && tc::conv_halo_supported(x.C, Co))
          return tc::conv_halo(x.p, x.pitch, x.coff, x.C, s.N, s.D, s.H, s.W, w.wcf[ci], Co, out.p, out.pitch, out.coff, 0, stats, st, nullptr, 1);
        return tc::conv(x.p, x.pitch, x.coff, x.C, s.N, s.D, s.H, s.W, w.wcf[ci], Co, ks, out.p, out.pitch, out.coff, 0, stats, st, 1);
      }
    }
    return simt_conv_fwd<T>(x, s, W, Co, ks, out, st);
  }
  // nb (bf16 engine only): fold the first pass of the backward of the norm that produced this conv's input into the epilogue
  int conv_dgrad(Cl<const T> dy, Sp s, const float* W, int Ci, int ks, Cl<T> dx, int acc, cudaStream_t st, const tc::HaloNormBwd* nb = nullptr) {
    if (nb && nb->done) *nb->done = false;
    if constexpr (kTC) {
      int ci = conv_index(W);
      if (ci >= 0 && tc::conv_supported(dy.C, Ci, dy.pitch, dy.coff, dx.pitch, dx.coff)) {
        B200_PROFD(st, "conv_dgrad k%d %d->%d @%d", ks, dy.C, Ci, s.D);
        if (ks == 3 && tc::conv_halo_supported(dy.C, Ci))
          return tc::conv_halo(dy.p, dy.pitch, dy.coff, dy.C, s.N, s.D, s.H, s.W, w.wcd[ci], Ci, dx.p, dx.pitch, dx.coff, acc, nullptr, st, nullptr, 0, nb);
        return tc::conv(dy.p, dy.pitch, dy.coff, dy.C, s.N, s.D, s.H, s.W, w.wcd[ci], Ci, ks, dx.p, dx.pitch, dx.coff, acc, nullptr, st);
      }
    }
    return simt_conv_dgrad<T>(dy, s, W, Ci, ks, dx, acc, st);
  }
  int conv_wgrad(Cl<const T> x, Cl<const T> dy, Sp s, int ks, float* dW, cudaStream_t st) {
    if constexpr (kTC) {
      if (tc::wgrad_supported(x.C, dy.C, x.pitch, x.coff, dy.pitch, dy.coff)) {
        B200_PROFD(st, "conv_wgrad k%d %dx%d @%d", ks, x.C, dy.C, s.D);
        if (tc::wgrad_halo_supported(x.C, dy.C, ks))
          return tc::conv_wgrad_halo(x.p, x.pitch, x.coff, x.C, dy.p, dy.pitch, dy.coff, dy.C, s.N, s.D, s.H, s.W, dW, st);
        return tc::conv_wgrad(x.p, x.pitch, x.coff, x.C, dy.p, dy.pitch, dy.coff, dy.C, s.N, s.D, s.H, s.W, ks, dW, st, w.wg_scratch, w.wg_scratch_bytes);
      }
    }
    return simt_conv_wgrad<T>(x, dy, s, ks, dW, st);
  }

  struct HeadArgs { const float* Wh; const float* bh; float* logits; };
  static bool edge_co_ok(int co) { return co == 8 || co == 16 || co == 32; }

  // raw != null: the block input is the NCDHW fp32 network input (encoder1) and conv1/conv3 run in the dedicated kernel.
  // head != null: the last normalise pass also emits the 1x1x1 head's logits (decoder2).
  int res_fwd(Cl<const T> x, int level, const float* W1, const float* W2, const float* W3, ResSave<T>& r, Cl<T> out, cudaStream_t st,
              const float* raw = nullptr, const HeadArgs* head = nullptr) {
    Sp s = sp(level); long Vs = V[level]; int Co = out.C;
    Cl<T> c1 = cl(w.c1tmp, Co, 0, Co), a1 = cl(r.a1, Co, 0, Co), c2 = cl(r.c2, Co, 0, Co), c3 = cl(r.c3, Co, 0, Co);
    bool done = false;
    if (raw) {
      B200_TRY(next_stat());
      double* st3 = w.stat_acc + (size_t)2 * c.B * 8 * c.fs;
      { B200_PROF("enc1_conv_fwd", st);
        dim3 g((unsigned)min(148L * 4, (Vs + 255) / 256), c.B);
        size_t sm = sizeof(float) * ((size_t)c.Cin * 27 * Co + (size_t)c.Cin * Co);
        if (s.W % 2 == 0 && !getenv("B200_ENC1_V1")) {   // two voxels x eight channels per thread
          dim3 g2((unsigned)min(148L * 8, (Vs / 2 * (Co / 8) + 255) / 256), c.B);
          if (Co == 8) conv_in_fwd2_kernel<T, 8><<<g2, 256, sm, st>>>(raw, W1, W3, c.Cin, s.D, s.H, s.W, (TR*)c1.p, (TR*)c3.p, w.stat_acc, st3);
          else if (Co == 16) conv_in_fwd2_kernel<T, 16><<<g2, 256, sm, st>>>(raw, W1, W3, c.Cin, s.D, s.H, s.W, (TR*)c1.p, (TR*)c3.p, w.stat_acc, st3);
          else conv_in_fwd2_kernel<T, 32><<<g2, 256, sm, st>>>(raw, W1, W3, c.Cin, s.D, s.H, s.W, (TR*)c1.p, (TR*)c3.p, w.stat_acc, st3);
        } else
        if (Co == 8) conv_in_fwd_kernel<T, 8><<<g, 256, sm, st>>>(raw, W1, W3, c.Cin, s.D, s.H, s.W, (TR*)c1.p, (TR*)c3.p, w.stat_acc, st3);
        else if (Co == 16) conv_in_fwd_kernel<T, 16><<<g, 256, sm, st>>>(raw, W1, W3, c.Cin, s.D, s.H, s.W, (TR*)c1.p, (TR*)c3.p, w.stat_acc, st3);
        else conv_in_fwd_kernel<T, 32><<<g, 256, sm, st>>>(raw, W1, W3, c.Cin, s.D, s.H, s.W, (TR*)c1.p, (TR*)c3.p, w.stat_acc, st3);
        B200_LAUNCH_CHECK();
        set_pending(r.mr1, w.stat_acc); set_pending(r.mr3, st3);
      }
      B200_TRY(in_apply(cl<const T>(c1.p, Co, 0, Co), r.mr1, nullptr, nullptr, a1, Vs, st));
      B200_TRY(conv_fwd(cl<const T>(a1.p, Co, 0, Co), s, W2, Co, 3, c2, w.stat_acc, &done, st));
      B200_TRY(in_stats(cl<const T>(c2.p, Co, 0, Co), Vs, r.mr2, st, done));
      B200_TRY(in_apply(cl<const T>(c2.p, Co, 0, Co), r.mr2, c3.p, r.mr3, out, Vs, st));
      return 0;
    }
    bool fused13 = false;   // conv1 (3^3) and conv3 (1^3) read the same x: one kernel, two accumulators, two stat sets
    if constexpr (kTC) {
      int i1 = conv_index(W1), i3 = conv_index(W3);
      if (i1 >= 0 && i3 >= 0 && tc::conv_supported(x.C, Co, x.pitch, x.coff, Co, 0) && tc::conv_halo_fused_supported(x.C, Co, 1)) {
        B200_PROFD(st, "conv_fwd k3+k1 %d->%d @%d", x.C, Co, s.D);
        B200_TRY(next_stat());
        double* st3 = w.stat_acc + (size_t)2 * c.B * 8 * c.fs;
        tc::HaloFused fu = {1, w.wcf[i3], c3.p, Co, 0, st3, nullptr, 0, 0};
        B200_TRY(tc::conv_halo(x.p, x.pitch, x.coff, x.C, s.N, s.D, s.H, s.W, w.wcf[i1], Co, c1.p, Co, 0, 0, w.stat_acc, st, &fu, 1));
        set_pending(r.mr1, w.stat_acc); set_pending(r.mr3, st3);
        fused13 = true;
      }
    }
    if (!fused13) {
      B200_TRY(conv_fwd(x, s, W1, Co, 3, c1, w.stat_acc, &done, st));
      B200_TRY(in_stats(cl<const T>(c1.p, Co, 0, Co), Vs, r.mr1, st, done));
    }
    B200_TRY(in_apply(cl<const T>(c1.p, Co, 0, Co), r.mr1, nullptr, nullptr, a1, Vs, st));
    B200_TRY(conv_fwd(cl<const T>(a1.p, Co, 0, Co), s, W2, Co, 3, c2, w.stat_acc, &done, st));
    B200_TRY(in_stats(cl<const T>(c2.p, Co, 0, Co), Vs, r.mr2, st, done));
    if (!fused13) {
      B200_TRY(conv_fwd(x, s, W3, Co, 1, c3, w.stat_acc, &done, st));
      B200_TRY(in_stats(cl<const T>(c3.p, Co, 0, Co), Vs, r.mr3, st, done));
    }
    if (head && out.pitch == Co && out.coff == 0) {
      B200_PROF("norm_head_fwd", st);
      dim3 g((unsigned)min(148L * 4, (Vs + 255) / 256), c.B);
      const double* a2 = take_pending(r.mr2); const double* a3 = take_pending(r.mr3); const double invV = 1.0 / (double)Vs;
      if (Co == 8) in_apply_head_kernel<T, 8><<<g, 256, 0, st>>>((const TR*)c2.p, r.mr2, (const TR*)c3.p, r.mr3, out.p, head->Wh, head->bh, c.ncls, Vs, head->logits, a2, a3, invV);
      else if (Co == 16) in_apply_head_kernel<T, 16><<<g, 256, 0, st>>>((const TR*)c2.p, r.mr2, (const TR*)c3.p, r.mr3, out.p, head->Wh, head->bh, c.ncls, Vs, head->logits, a2, a3, invV);
      else in_apply_head_kernel<T, 32><<<g, 256, 0, st>>>((const TR*)c2.p, r.mr2, (const TR*)c3.p, r.mr3, out.p, head->Wh, head->bh, c.ncls, Vs, head->logits, a2, a3, invV);
      B200_LAUNCH_CHECK();
      return 0;
    }
    B200_TRY(in_apply(cl<const T>(c2.p, Co, 0, Co), r.mr2, c3.p, r.mr3, out, Vs, st));
    return 0;
  }
  // dOut: gradient wrt block output; out: the block's forward output.  Writes dW1..3 (if non-null) and, if dx.p, the
  // input gradient (dx = dgrad3(dc1) + dgrad1(dc3)).
  int res_bwd(Cl<const T> x, int level, const float* W1, const float* W2, const float* W3, ResSave<T>& r, Cl<const T> out,
              Cl<const T> dOut, float* dW1, float* dW2, float* dW3, Cl<T> dx, cudaStream_t st, const float* raw = nullptr) {
    constexpr int VN = Vec16<T>::N;
    Sp s = sp(level); long Vs = V[level]; int Co = out.C, Ci = x.C; int B = c.B;
    ClView pv{Co, 0};
    // data-parallel overlap: the weight gradients of the conv stacks are issued AFTER the ViT backward (flush_deferred), so that the
    // 340 MB of ViT gradients are reduced behind them; their inputs dc1..3 then live in per-block buffers instead of the shared ones
    const int blk = (int)(&r - w.rs);
    T* const pdc1 = defer_wg ? w.dcs[blk][0] : w.dc1; T* const pdc2 = defer_wg ? w.dcs[blk][1] : w.dc2; T* const pdc3 = defer_wg ? w.dcs[blk][2] : w.dc3;
    size_t red_smem = 256 * 3 * VN * sizeof(float), cst_smem = 7 * (size_t)Co * sizeof(float);
    // the sign of the block output is recomputed from c2, c3 instead of reading `out` again (B200_INBWD_ACT=1 reads it)
    static const bool read_act = getenv("B200_INBWD_ACT") != nullptr;
    const T* actp = read_act ? out.p : (const T*)nullptr;
    static const int grmul = getenv("B200_INBWD_GRID") ? atoi(getenv("B200_INBWD_GRID")) : 1;
    dim3 gr(in_grid_x(Vs, Co / VN) * grmul, B), ga(in_grid_x(Vs, Co / VN) * 2, B);
    // final lrelu + two norms
    { B200_PROF("instnorm_bwd", st);
    B200_TRY(next_bwd());
    if (read_act) {
      B200_CUDA(launch_pdl(in_bwd_reduce_kernel<T, true, false>, dim3(gr), dim3(256), red_smem, st, dOut.p, ClView{dOut.pitch, dOut.coff}, actp, ClView{out.pitch, out.coff},
                                                        (const TR*)r.c2, pv, (const TR*)r.c3, pv, Co, Vs, w.bwd_acc, (const float*)r.mr2, (const float*)r.mr3));
      B200_LAUNCH_CHECK();
      B200_CUDA(launch_pdl(in_bwd_apply_kernel<T, true, false>, dim3(ga), dim3(256), cst_smem, st, dOut.p, ClView{dOut.pitch, dOut.coff}, actp, ClView{out.pitch, out.coff},
                                                 (const TR*)r.c2, pv, r.mr2, (const TR*)r.c3, pv, r.mr3, Co, Vs, w.bwd_acc, pdc2, pv, pdc3, pv));
    } else {
      B200_CUDA(launch_pdl(in_bwd_reduce_kernel<T, true, true>, dim3(gr), dim3(256), red_smem, st, dOut.p, ClView{dOut.pitch, dOut.coff}, actp, ClView{out.pitch, out.coff},
                                                        (const TR*)r.c2, pv, (const TR*)r.c3, pv, Co, Vs, w.bwd_acc, (const float*)r.mr2, (const float*)r.mr3));
      B200_LAUNCH_CHECK();
      B200_CUDA(launch_pdl(in_bwd_apply_kernel<T, true, true>, dim3(ga), dim3(256), cst_smem, st, dOut.p, ClView{dOut.pitch, dOut.coff}, actp, ClView{out.pitch, out.coff},
                                                 (const TR*)r.c2, pv, r.mr2, (const TR*)r.c3, pv, r.mr3, Co, Vs, w.bwd_acc, pdc2, pv, pdc3, pv));
    }
    B200_LAUNCH_CHECK(); }
    Cl<const T> dc2 = cl<const T>(pdc2, Co, 0, Co), dc3 = cl<const T>(pdc3, Co, 0, Co), a1 = cl<const T>(r.a1, Co, 0, Co);
    // conv2
    if (dW2) B200_TRY(run_or_defer([=](cudaStream_t st) -> int { B200_TRY(zero_grad(dW2, (size_t)Co * Co * 27, st)); return conv_wgrad(a1, dc2, s, 3, dW2, st); }, st));
    // lrelu + norm1: the reduction pass (sum g, sum g*n) rides on the dgrad epilogue where the kernel supports it (tc_conv_halo48.cuh)
    B200_TRY(next_bwd());
    bool folded = false;
    if constexpr (kTC) {
      // MEASURED SLOWER on B200 (configs[1]: 5.34 ms/step folded vs 5.21 ms with the separate reduction pass): the dgrad kernel is tensor-bound
      // with its epilogue hidden behind the next tile's MMAs, and the extra 32-byte global load per row in the epilogue un-hides it (the
      // saved 2 x 30 us of in_bwd_reduce cost 2 x 95 us in the convolution).  Kept as an opt-in experiment (B200_NORM_FOLD=1) with its
      // op-level test (tests/test_gpu_tcconv.py).
      static const bool fold = getenv("B200_NORM_FOLD") != nullptr;
      tc::HaloNormBwd nb = {reinterpret_cast<const bf16*>(r.a1), Co, 0, w.bwd_acc, &folded};
      B200_TRY(conv_dgrad(dc2, s, W2, Co, 3, cl(w.da1, Co, 0, Co), 0, st, fold ? &nb : nullptr));
    } else {
      B200_TRY(conv_dgrad(dc2, s, W2, Co, 3, cl(w.da1, Co, 0, Co), 0, st));
    }
    { B200_PROF("instnorm_bwd", st);
    if (!folded) {
    B200_CUDA(launch_pdl(in_bwd_reduce_kernel<T, false>, dim3(gr), dim3(256), red_smem, st, w.da1, pv, r.a1, pv, (const TR*)nullptr, pv, (const TR*)nullptr, pv, Co, Vs, w.bwd_acc, (const float*)nullptr, (const float*)nullptr));
    B200_LAUNCH_CHECK(); }
    B200_CUDA(launch_pdl(in_bwd_apply_kernel<T, false>, dim3(ga), dim3(256), cst_smem, st, w.da1, pv, r.a1, pv, (const TR*)nullptr, pv, r.mr1, (const TR*)nullptr, pv, nullptr, Co, Vs, w.bwd_acc,
                                               pdc1, pv, nullptr, pv));
    B200_LAUNCH_CHECK(); }
    Cl<const T> dc1 = cl<const T>(pdc1, Co, 0, Co);
    if (raw) {   // encoder1: both weight gradients from the raw fp32 input in one dedicated kernel
      B200_CHECK(dW1 && dW3, "encoder1 weight gradients are produced together");
      return run_or_defer([=](cudaStream_t st) -> int {
      B200_PROF("enc1_conv_wgrad", st);
      B200_TRY(zero_grad(dW1, (size_t)Co * Ci * 27, st));
      B200_TRY(zero_grad(dW3, (size_t)Co * Ci, st));
      long chunk = 4096;
      dim3 g((unsigned)((Vs + chunk - 1) / chunk), B, Ci);
      if (s.W % 2 == 0 && !getenv("B200_ENC1_V1")) {   // two voxels x eight channels x nine taps per thread
        const long ppb = 4096;
        dim3 g2((unsigned)((Vs / 2 + ppb - 1) / ppb), B, (unsigned)(3 * (Co / 8) * Ci));
        if (Co == 8) conv_in_wgrad2_kernel<T, 8><<<g2, 256, 0, st>>>(raw, pdc1, pdc3, Ci, s.D, s.H, s.W, ppb, dW1, dW3);
        else if (Co == 16) conv_in_wgrad2_kernel<T, 16><<<g2, 256, 0, st>>>(raw, pdc1, pdc3, Ci, s.D, s.H, s.W, ppb, dW1, dW3);
        else conv_in_wgrad2_kernel<T, 32><<<g2, 256, 0, st>>>(raw, pdc1, pdc3, Ci, s.D, s.H, s.W, ppb, dW1, dW3);
      } else
      if (Co == 8) conv_in_wgrad_kernel<T, 8><<<g, 256, 0, st>>>(raw, pdc1, pdc3, Ci, s.D, s.H, s.W, chunk, dW1, dW3);
      else if (Co == 16) conv_in_wgrad_kernel<T, 16><<<g, 256, 0, st>>>(raw, pdc1, pdc3, Ci, s.D, s.H, s.W, chunk, dW1, dW3);
      else conv_in_wgrad_kernel<T, 32><<<g, 256, 0, st>>>(raw, pdc1, pdc3, Ci, s.D, s.H, s.W, chunk, dW1, dW3);
      B200_LAUNCH_CHECK();
      return 0; }, st);
    }
    B200_TRY(run_or_defer([=](cudaStream_t st) -> int {
    bool wg_fused = false;
    if constexpr (kTC) {
      // conv1 (3^3) and conv3 (1^3) read the same x: one halo weight-gradient launch with a tenth accumulator for the 1^3 tap
      static const bool off = getenv("B200_NO_FUSED_WGRAD_1X1") != nullptr;
      if (!off && dW1 && dW3 && tc::wgrad_supported(x.C, dc1.C, x.pitch, x.coff, dc1.pitch, dc1.coff) && tc::wgrad_halo_supported(x.C, dc1.C, 3) &&
          dc3.C == dc1.C && (dc3.pitch * 2) % 16 == 0 && (dc3.coff * 2) % 16 == 0) {
        B200_TRY(zero_grad(dW1, (size_t)Co * Ci * 27, st));
        B200_TRY(zero_grad(dW3, (size_t)Co * Ci, st));
        B200_PROFD(st, "conv_wgrad k3+k1 %dx%d @%d", x.C, dc1.C, s.D);
        B200_TRY(tc::conv_wgrad_halo(x.p, x.pitch, x.coff, x.C, dc1.p, dc1.pitch, dc1.coff, dc1.C, s.N, s.D, s.H, s.W, dW1, st,
                                     dc3.p, dc3.pitch, dc3.coff, dW3));
        wg_fused = true;
      }
    }
    if (!wg_fused) {
      if (dW1) { B200_TRY(zero_grad(dW1, (size_t)Co * Ci * 27, st)); B200_TRY(conv_wgrad(x, dc1, s, 3, dW1, st)); }
      if (dW3) { B200_TRY(zero_grad(dW3, (size_t)Co * Ci, st)); B200_TRY(conv_wgrad(x, dc3, s, 1, dW3, st)); }
    }
    return 0; }, st));
    if (dx.p) {
      if constexpr (kTC) {   // dx = dgrad3x3(dc1) + dgrad1x1(dc3) in one kernel (second input tile, same accumulator)
        int i1 = conv_index(W1), i3 = conv_index(W3);
        if (i1 >= 0 && i3 >= 0 && tc::conv_supported(Co, Ci, Co, 0, dx.pitch, dx.coff) && tc::conv_halo_fused_supported(Co, Ci, 2)) {
          B200_PROFD(st, "conv_dgrad k3+k1 %d->%d @%d", Co, Ci, s.D);
          tc::HaloFused fu = {2, w.wcd[i3], nullptr, 0, 0, nullptr, dc3.p, dc3.pitch, dc3.coff};
          return tc::conv_halo(dc1.p, dc1.pitch, dc1.coff, Co, s.N, s.D, s.H, s.W, w.wcd[i1], Ci, dx.p, dx.pitch, dx.coff, 0, nullptr, st, &fu);
        }
      }
      B200_TRY(conv_dgrad(dc1, s, W1, Ci, 3, dx, 0, st));
      B200_TRY(conv_dgrad(dc3, s, W3, Ci, 1, dx, 1, st));
    }
    return 0;
  }

  // ------------------------------------------------------------ forward
  int forward(const float* const* P, const float* x_in, char* ws, float* enc4_out, float* logits_out, int flags, cudaStream_t st) {
    layout(ws, false);
    if (kTC) { B200_CHECK(packed_base, "bf16 mode needs the packed-weight buffer (b200_unetr_set_packed_weights)"); layout_packed(packed_base); }
    no_backward = (flags & FLAG_NO_BACKWARD) != 0;
    for (auto& p : pend_stat) p = PendStat{nullptr, nullptr};   // nothing carries over from a forward that ended early
    B200_CUDA(cudaMemsetAsync(w.stat_pool, 0, sizeof(double) * kStatSlots * 4 * c.B * 8 * c.fs, st));
    B200_PROFC_BEGIN("F1 pack+patch", st);
    int B = c.B, fs = c.fs;
    if (!(flags & FLAG_WEIGHTS_PACKED)) B200_TRY(pack_weights(P, st));
    cur_params = P;
    // --- patch embedding (a5): tokens = rows(x) W^T + b + pos
    {
      B200_PROF("patch_fwd", st);
      EpPatch ep = {w.x0, H, L, P[P_PATCH_B], P[P_POS]};
      if constexpr (kTC) {
        long tot = (long)M * 4096 * c.Cin;
        // S2 % 16 == 0 (img_size is a multiple of the patch size), so 8-voxel groups of a patch line are 32-byte aligned in an fp32 volume
        if ((c.conv_patch || c.Cin == 1) && tot / 8 < (1L << 30) && ((uintptr_t)x_in & 15) == 0)
          patch_gather8_kernel<<<(unsigned)min(148L * 8, (tot / 8 + 255) / 256), 256, 0, st>>>(x_in, w.apatch, c.Cin, c.S0, c.S1, c.S2, g0, g1, g2, (int)(tot / 8));
        else
          patch_gather_kernel<<<(unsigned)min(148L * 8, (tot + 255) / 256), 256, 0, st>>>(x_in, w.apatch, c.Cin, c.S0, c.S1, c.S2, g0, g1, g2, c.conv_patch, tot);
        B200_LAUNCH_CHECK();
        B200_TRY(tc::gemm(tc::operand(w.apatch, 4096L * c.Cin, 1), tc::operand(w.wpatch, 4096L * c.Cin, 1), ep, M, H, 4096 * c.Cin, 1, 1, st));
      } else {
        RowIsOuter<PatchGather, false> al; al.g = {x_in, c.Cin, c.S0, c.S1, c.S2, g0, g1, g2, c.conv_patch};
        B200_TRY(launch_contract(al, ld2<float, false>(P[P_PATCH_W], 4096L * c.Cin, 1), ep, M, H, 4096 * c.Cin, 1, 1, st));
      }
    }
    B200_PROFC_END(st); B200_PROFC_BEGIN("F2 vit", st);
    // --- transformer blocks (a6-a8)
    float scale = 1.0f / sqrtf((float)dh);
    SplitSum pend; memset(&pend, 0, sizeof(pend));   // split-K partials of the previous N = hidden GEMM, consumed by the next LayerNorm
    int hsT_done = 0;
    for (int i = 0; i < 12; ++i) {
      const float* const* bp = P + P_BLK0 + i * B_COUNT;
      const float* xin = i ? w.hs[i - 1] : w.x0;
      // hidden states 3 / 6 / 9 (= hs[i - 1] here) feed the encoders as T: the LayerNorm that forms them from the split-K partials
      // stores that copy too
      if (pend.nsplit && (i == 4 || i == 7 || i == 10)) { pend.xsum_cast = w.hsT[(i - 4) / 3]; hsT_done |= 1 << ((i - 4) / 3); }
      B200_TRY(launch_layernorm_fwd<T>(xin, bp[B_LN1_W], bp[B_LN1_B], w.ln1[i], w.ln1s[i], M, H, st, pend.nsplit ? &pend : nullptr));
      pend.nsplit = 0; pend.xsum_cast = nullptr;
      B200_TRY(linear_fwd<T>(w.ln1[i], H, bp[B_QKV_W], w.wqkv[i], M, 3 * H, H, ep_plain<T>(w.qkv[i], 3 * H), st));
      B200_TRY(attention_fwd(i, scale, st));
      if (can_split(M, H, H)) {
        B200_TRY(linear_fwd_split(w.att[i], H, w.wproj[i], M, H, H, bp[B_PROJ_B], xin, w.x1[i], &pend, st));
      } else {
        EpStore<float> ep = ep_plain<float>(w.x1[i], H); ep.bias = bp[B_PROJ_B]; ep.resid = xin; ep.ldr = H;
        B200_TRY(linear_fwd<float>(w.att[i], H, bp[B_PROJ_W], w.wproj[i], M, H, H, ep, st));
      }
      B200_TRY(launch_layernorm_fwd<T>(w.x1[i], bp[B_LN2_W], bp[B_LN2_B], w.ln2[i], w.ln2s[i], M, H, st, pend.nsplit ? &pend : nullptr));
      pend.nsplit = 0;
      { EpStore<T> ep = ep_plain<T>(w.h[i], F); ep.bias = bp[B_FC1_B];
        // training: u[i] receives gelu'(pre-activation) (all the backward needs of it); inference: nothing is kept
        if (no_backward) { ep.act = ACT_GELU; ep.preact = nullptr; } else { ep.act = ACT_GELU_SAVE; ep.preact = w.u[i]; }
        B200_TRY(linear_fwd<T>(w.ln2[i], H, bp[B_FC1_W], w.wfc1[i], M, F, H, ep, st)); }
      if (can_split(M, H, F)) {
        B200_TRY(linear_fwd_split(w.h[i], F, w.wfc2[i], M, H, F, bp[B_FC2_B], w.x1[i], w.hs[i], &pend, st));
      } else {
        EpStore<float> ep = ep_plain<float>(w.hs[i], H); ep.bias = bp[B_FC2_B]; ep.resid = w.x1[i]; ep.ldr = H;
        B200_TRY(linear_fwd<float>(w.h[i], F, bp[B_FC2_W], w.wfc2[i], M, H, F, ep, st));
      }
    }
    B200_TRY(launch_layernorm_fwd<T>(w.hs[11], P[P_NORM_W], P[P_NORM_B], w.vit_out, w.lnfs, M, H, st, pend.nsplit ? &pend : nullptr));
    for (int k = 0; k < 3; ++k) if (!(hsT_done >> k & 1)) B200_TRY(launch_cast<float, T>(w.hs[3 + 3 * k], w.hsT[k], (long)M * H, st));

    B200_PROFC_END(st); B200_PROFC_BEGIN("F3 encoders", st);
    // --- encoder1 on the input volume (a9) -> upper half of concat2
    const bool edge = edge_co_ok(fs);   // dedicated small-channel kernels (fp32 input, fused head) available for this feature_size
    if (!edge) B200_TRY(launch_layout<T>(x_in, nullptr, w.xcl, B, c.Cin, V[0], c.Cin, 0, 0, 0, st));
    B200_TRY(res_fwd(cl<const T>(w.xcl, c.Cin, 0, c.Cin), 0, P[P_E1_C1], P[P_E1_C2], P[P_E1_C3], w.rs[0], cl(w.cat2, 2 * fs, fs, fs), st,
                     edge ? x_in : nullptr));
    // --- encoder2..4: transposed-conv pyramids from hidden states 3/6/9 (a10)
    B200_TRY(convT_fwd(w.hsT[0], H, H, 4, P[P_E2_T0], cl(w.e2a, 2 * fs, 0, 2 * fs), st));
    B200_TRY(convT_fwd(w.e2a, 2 * fs, 2 * fs, 3, P[P_E2_T1], cl(w.e2b, 2 * fs, 0, 2 * fs), st));
    B200_TRY(convT_fwd(w.e2b, 2 * fs, 2 * fs, 2, P[P_E2_T2], cl(w.cat3, 4 * fs, 2 * fs, 2 * fs), st));
    B200_TRY(convT_fwd(w.hsT[1], H, H, 4, P[P_E3_T0], cl(w.e3a, 4 * fs, 0, 4 * fs), st));
    B200_TRY(convT_fwd(w.e3a, 4 * fs, 4 * fs, 3, P[P_E3_T1], cl(w.cat4, 8 * fs, 4 * fs, 4 * fs), st));
    B200_TRY(convT_fwd(w.hsT[2], H, H, 4, P[P_E4_T0], cl(w.cat5, 16 * fs, 8 * fs, 8 * fs), st));
    if (enc4_out) B200_TRY(launch_layout<T>(nullptr, enc4_out, w.cat5, B, 8 * fs, V[3], 16 * fs, 8 * fs, 1, 0, st));
    B200_PROFC_END(st); B200_PROFC_BEGIN("F4 decoders+head", st);
    // --- decoder5..2 (a11)
    B200_TRY(convT_fwd(w.vit_out, H, H, 4, P[P_D5_T], cl(w.cat5, 16 * fs, 0, 8 * fs), st));
    B200_TRY(res_fwd(cl<const T>(w.cat5, 16 * fs, 0, 16 * fs), 3, P[P_D5_C1], P[P_D5_C2], P[P_D5_C3], w.rs[1], cl(w.d3, 8 * fs, 0, 8 * fs), st));
    B200_TRY(convT_fwd(w.d3, 8 * fs, 8 * fs, 3, P[P_D4_T], cl(w.cat4, 8 * fs, 0, 4 * fs), st));
    B200_TRY(res_fwd(cl<const T>(w.cat4, 8 * fs, 0, 8 * fs), 2, P[P_D4_C1], P[P_D4_C2], P[P_D4_C3], w.rs[2], cl(w.d2, 4 * fs, 0, 4 * fs), st));
    B200_TRY(convT_fwd(w.d2, 4 * fs, 4 * fs, 2, P[P_D3_T], cl(w.cat3, 4 * fs, 0, 2 * fs), st));
    B200_TRY(res_fwd(cl<const T>(w.cat3, 4 * fs, 0, 4 * fs), 1, P[P_D3_C1], P[P_D3_C2], P[P_D3_C3], w.rs[3], cl(w.d1, 2 * fs, 0, 2 * fs), st));
    B200_TRY(convT_fwd(w.d1, 2 * fs, 2 * fs, 1, P[P_D2_T], cl(w.cat2, 2 * fs, 0, fs), st));
    HeadArgs head = {P[P_OUT_W], P[P_OUT_B], logits_out};
    const bool fused_head = edge && logits_out;
    B200_TRY(res_fwd(cl<const T>(w.cat2, 2 * fs, 0, 2 * fs), 0, P[P_D2_C1], P[P_D2_C2], P[P_D2_C3], w.rs[4], cl(w.d0, fs, 0, fs), st, nullptr,
                     fused_head ? &head : nullptr));
    // --- 1x1x1 head with bias, NCDHW fp32 logits (a12) -- normally fused into the pass above
    if (logits_out && !fused_head) {
      B200_PROF("head_fwd", st);
      EpHeadNcdhw ep = {logits_out, c.ncls, V[0], P[P_OUT_B]};
      B200_TRY(launch_contract(ld2<T, false>(w.d0, fs, 1), ld2<float, false>(P[P_OUT_W], fs, 1), ep, (int)(B * V[0]), c.ncls, fs, 1, 1, st));
    }
    B200_PROFC_END(st);
    return 0;
  }

  int convT_fwd(const T* x, long ldx, int Ci, int in_level, const float* W, Cl<T> out, cudaStream_t st) {
    if constexpr (kTC) {
      if (cur_params) {
        const bool tapm = out.C % 16 == 0 && out.pitch % 8 == 0 && out.coff % 8 == 0;
        const bf16* Wb = convT_packed(cur_params, W, tapm);
        if (Wb && tapm) {   // tap-major weight: 16 consecutive GEMM columns = 16 channels of one output voxel (32-byte stores)
          B200_PROFD(st, "convT_fwd %d->%d @%d", Ci, out.C, sp(in_level).D);
          Sp s = sp(in_level);
          EpConvTScatterTap<T> ep = {out.p, s.D, s.H, s.W, out.pitch, out.coff, out.C};
          return tc::gemm(tc::operand(x, ldx, 1), tc::operand(Wb, 1, (long)out.C * 8), ep, (int)s.rows(), out.C * 8, Ci, 1, 1, st);
        }
        if (Wb) {   // [rows, Ci] x W[Ci][Co*8] (MN-major B), scatter epilogue writes straight into the concat buffer
          B200_PROFD(st, "convT_fwd %d->%d @%d", Ci, out.C, sp(in_level).D);
          Sp s = sp(in_level);
          EpConvTScatter<T> ep = {out.p, s.D, s.H, s.W, out.pitch, out.coff};
          return tc::gemm(tc::operand(x, ldx, 1), tc::operand(Wb, 1, (long)out.C * 8), ep, (int)s.rows(), out.C * 8, Ci, 1, 1, st);
        }
      }
    }
    return simt_convT_fwd<T, T>(x, ldx, Ci, sp(in_level), W, out, st);
  }
  const float* const* cur_params = nullptr;
  bool no_backward = false;     // FLAG_NO_BACKWARD of the current forward: buffers only the backward reads are not written
  // Gradient-ready events (ExecIface::set_grad_events): event 0 = conv stacks + head, event k >= 1 = the k-th group of transformer
  // blocks from the top (the first also covers vit.norm, the last also the patch embedding).  4 / 7 / 13 events = groups of 4 / 2 / 1
  // blocks; the deferred parameter-gradient launches (vit_param_grads) use the same grouping.
  cudaEvent_t grad_ev[13]; int n_grad_ev = 0;
  // With gradient events set (data-parallel overlap) the conv-stack weight gradients only feed the optimizer, like the ViT parameter
  // gradients: they are queued here and launched after the ViT backward, and event 0 is recorded LAST (the host side reduces the
  // conv range last: unetr.py).  FLAG_INPLACE_WGRADS keeps the in-place order.
  bool defer_wg = false;
  std::vector<std::function<int(cudaStream_t)>> deferred;
  template <class F> int run_or_defer(F&& job, cudaStream_t st) {
    if (!defer_wg) return job(st);
    deferred.emplace_back(std::forward<F>(job));
    return 0;
  }
  int flush_deferred(cudaStream_t st) {
    int rc = 0;
    for (auto& j : deferred) { rc = j(st); if (rc) break; }
    deferred.clear();
    return rc;
  }
  int vit_group() const { return n_grad_ev == 13 ? 1 : n_grad_ev == 7 ? 2 : 4; }
  void mark_grads(int k, cudaStream_t st) { if (k < n_grad_ev) cudaEventRecord(grad_ev[k], st); }

  // softmax(Q K^T * scale) V per (batch, head); scores fp32, probabilities T (kept for backward)
  int attention_fwd(int i, float scale, cudaStream_t st) {
    B200_PROF("attention_fwd", st);
    const T* qkv = w.qkv[i];
    long sQb = (long)L * 3 * H, sPb = (long)nh * L * Lp, sPh = (long)L * Lp;
    int BH = c.B * nh;
    bool fused_sm = false;
    if constexpr (kTC) {
      // one kernel per layer: S = Q K^T in TMEM, softmax per TMEM lane, P through shared memory into the P V MMA (tc_attention.cuh)
      if (tc::attention_fused_supported(L, Lp, H, nh))
        return tc::attention_fused_fwd(qkv, no_backward ? nullptr : w.P[i], w.att[i], c.B, nh, L, Lp, H, scale, st);
    }
    if constexpr (kTC) {
      // scores stay in TMEM: softmax in the QK^T epilogue.  Measured SLOWER at L = 216 (0.30 -> 0.38 ms per 12 layers): one 224-wide tile
      // per (head, row block) leaves 48 CTAs with a serial two-pass epilogue; opt-in experiment only
      if (L <= 256 && Lp <= 256 && getenv("B200_FUSED_SOFTMAX")) {
        EpSoftmaxRow<T> ep = {w.P[i], sPb, sPh, nh, Lp, scale};
        B200_TRY(tc::gemm(tc::operand(qkv, 3 * H, 1, sQb, dh), tc::operand(qkv + H, 3 * H, 1, sQb, dh), ep, L, L, dh, c.B, nh, st));
        fused_sm = true;
      }
    }
    if (!fused_sm) {
    { EpStore<float> ep = ep_plain<float>(w.S, Lp); ep.sb0 = sPb; ep.sb1 = sPh; ep.nb1 = nh;
      if constexpr (kTC) B200_TRY(tc::gemm(tc::operand(qkv, 3 * H, 1, sQb, dh), tc::operand(qkv + H, 3 * H, 1, sQb, dh), ep, L, L, dh, c.B, nh, st));
      else B200_TRY(launch_contract(ld4<T, false>(qkv, 3 * H, 1, sQb, dh, nh), ld4<T, false>(qkv + H, 3 * H, 1, sQb, dh, nh), ep, L, L, dh, BH, 1, st)); }
    B200_CUDA(launch_pdl(softmax_fwd_kernel<T>, dim3(cdiv((long)BH * L, 8)), dim3(256), 0, st, w.S, w.P[i], (long)BH * L, L, Lp, scale));
    B200_LAUNCH_CHECK();
    }
    { EpStore<T> ep = ep_plain<T>(w.att[i], H); ep.sb0 = (long)L * H; ep.sb1 = dh; ep.nb1 = nh;
      if constexpr (kTC) B200_TRY(tc::gemm(tc::operand(w.P[i], Lp, 1, sPb, sPh), tc::operand(qkv + 2 * H, 1, 3 * H, sQb, dh), ep, L, dh, L, c.B, nh, st));
      else B200_TRY(launch_contract(ld4<T, false>(w.P[i], Lp, 1, sPb, sPh, nh), ld4<T, true>(qkv + 2 * H, 1, 3 * H, sQb, dh, nh), ep, L, dh, L, BH, 1, st)); }
    return 0;
  }
  // d(att) -> dqkv
  int attention_bwd(int i, float scale, cudaStream_t st) {
    B200_PROF("attention_bwd", st);
    const T* qkv = w.qkv[i];
    long sQb = (long)L * 3 * H, sPb = (long)nh * L * Lp, sPh = (long)L * Lp, sOb = (long)L * H;
    int BH = c.B * nh;
    if constexpr (kTC) {
      // query-row half in one kernel (tc_attention.cuh, BWD = 1): dP = dO V^T in TMEM -> dS (in place of the TMA-loaded P block in
      // smem) -> dQ = dS K; the key-row half stays two batched GEMMs over P^T and dS^T
      static const bool off = getenv("B200_NO_FUSED_ATTENTION_BWD") != nullptr;
      if (!off && tc::attention_fused_supported(L, Lp, H, nh)) {
        B200_TRY(tc::attention_fused_bwd_dq(qkv, w.P[i], w.datt, w.dS, w.dqkv[i], c.B, nh, L, Lp, H, scale, st));
        static const bool kv_off = getenv("B200_NO_FUSED_ATTENTION_KV") != nullptr;
        if (!kv_off) return tc::attention_fused_bwd_kv(qkv, w.P[i], w.dS, w.datt, w.dqkv[i], c.B, nh, L, Lp, H, st);   // dV, dK in one launch
        { EpStore<T> ep = ep_plain<T>(w.dqkv[i] + 2 * H, 3 * H); ep.sb0 = sQb; ep.sb1 = dh; ep.nb1 = nh;      // dV = P^T dO
          B200_TRY(tc::gemm(tc::operand(w.P[i], 1, Lp, sPb, sPh), tc::operand(w.datt, 1, H, sOb, dh), ep, L, dh, L, c.B, nh, st)); }
        { EpStore<T> ep = ep_plain<T>(w.dqkv[i] + H, 3 * H); ep.sb0 = sQb; ep.sb1 = dh; ep.nb1 = nh;          // dK = dS^T Q
          B200_TRY(tc::gemm(tc::operand(w.dS, 1, Lp, sPb, sPh), tc::operand(qkv, 1, 3 * H, sQb, dh), ep, L, dh, L, c.B, nh, st)); }
        return 0;
      }
    }
    // dP = dO V^T ; dS = P * (dP - rowsum(dP * P)) * scale  -- in the GEMM epilogue when a row fits one tile
    bool fused_sm = false;
    if constexpr (kTC) {
      if (L <= 256 && Lp <= 256 && getenv("B200_FUSED_SOFTMAX")) {   // see attention_fwd: slower at L = 216, opt-in
        EpSoftmaxBwdRow<T> ep = {w.P[i], w.dS, sPb, sPh, nh, Lp, scale};
        B200_TRY(tc::gemm(tc::operand(w.datt, H, 1, sOb, dh), tc::operand(qkv + 2 * H, 3 * H, 1, sQb, dh), ep, L, L, dh, c.B, nh, st));
        fused_sm = true;
      }
    }
    if (!fused_sm) {
      EpStore<float> ep = ep_plain<float>(w.dP, Lp); ep.sb0 = sPb; ep.sb1 = sPh; ep.nb1 = nh;
      if constexpr (kTC) B200_TRY(tc::gemm(tc::operand(w.datt, H, 1, sOb, dh), tc::operand(qkv + 2 * H, 3 * H, 1, sQb, dh), ep, L, L, dh, c.B, nh, st));
      else B200_TRY(launch_contract(ld4<T, false>(w.datt, H, 1, sOb, dh, nh), ld4<T, false>(qkv + 2 * H, 3 * H, 1, sQb, dh, nh), ep, L, L, dh, BH, 1, st));
    }
    // dV = P^T dO
    { EpStore<T> ep = ep_plain<T>(w.dqkv[i] + 2 * H, 3 * H); ep.sb0 = sQb; ep.sb1 = dh; ep.nb1 = nh;
      if constexpr (kTC) B200_TRY(tc::gemm(tc::operand(w.P[i], 1, Lp, sPb, sPh), tc::operand(w.datt, 1, H, sOb, dh), ep, L, dh, L, c.B, nh, st));
      else B200_TRY(launch_contract(ld4<T, true>(w.P[i], 1, Lp, sPb, sPh, nh), ld4<T, true>(w.datt, 1, H, sOb, dh, nh), ep, L, dh, L, BH, 1, st)); }
    if (!fused_sm) {
      B200_CUDA(launch_pdl(softmax_bwd_kernel<T>, dim3(cdiv((long)BH * L, 8)), dim3(256), 0, st, w.P[i], w.dP, w.dS, (long)BH * L, L, Lp, scale));
      B200_LAUNCH_CHECK();
    }
    // dQ = dS K ; dK = dS^T Q
    { EpStore<T> ep = ep_plain<T>(w.dqkv[i], 3 * H); ep.sb0 = sQb; ep.sb1 = dh; ep.nb1 = nh;
      if constexpr (kTC) B200_TRY(tc::gemm(tc::operand(w.dS, Lp, 1, sPb, sPh), tc::operand(qkv + H, 1, 3 * H, sQb, dh), ep, L, dh, L, c.B, nh, st));
      else B200_TRY(launch_contract(ld4<T, false>(w.dS, Lp, 1, sPb, sPh, nh), ld4<T, true>(qkv + H, 1, 3 * H, sQb, dh, nh), ep, L, dh, L, BH, 1, st)); }
    { EpStore<T> ep = ep_plain<T>(w.dqkv[i] + H, 3 * H); ep.sb0 = sQb; ep.sb1 = dh; ep.nb1 = nh;
      if constexpr (kTC) B200_TRY(tc::gemm(tc::operand(w.dS, 1, Lp, sPb, sPh), tc::operand(qkv, 1, 3 * H, sQb, dh), ep, L, dh, L, c.B, nh, st));
      else B200_TRY(launch_contract(ld4<T, true>(w.dS, 1, Lp, sPb, sPh, nh), ld4<T, true>(qkv, 1, 3 * H, sQb, dh, nh), ep, L, dh, L, BH, 1, st)); }
    return 0;
  }

  // Parameter gradients of transformer blocks lo..hi from the per-block operands the critical path left behind:
  //   weights   dW = dY^T X for {fc2, fc1, out_proj, qkv}: ONE grouped tcgen05 launch (tc_gemm_grouped.cuh) in bf16 mode
  //   biases    column sums of dY: one launch;   LayerNorm weight / bias: one launch.
  int vit_param_grads(float* const* G, int lo, int hi, cudaStream_t st) {
    if (hi < lo) return 0;
    tc::GroupItem items[4 * 12]; int ni = 0; double flops = 0;
    ColsumJobs cs; cs.count = 0;
    LnParamJobs ln; ln.count = 0; ln.M = M; ln.H = H;
    for (int i = hi; i >= lo; --i) {
      float* const* bg = G + P_BLK0 + i * B_COUNT;
      const float* xin = i ? w.hs[i - 1] : w.x0;
      struct { float* dW; const T* dY; int N; const T* X; int K; } wg[4] = {
          {bg[B_FC2_W], w.dyo[i + 1], H, w.h[i], F}, {bg[B_FC1_W], w.dh[i], F, w.ln2[i], H},
          {bg[B_PROJ_W], w.dy1[i], H, w.att[i], H}, {bg[B_QKV_W], w.dqkv[i], 3 * H, w.ln1[i], H}};
      for (auto& g : wg) {
        if (!g.dW) continue;
        if constexpr (kTC) {
          tc::GroupItem& it = items[ni++];
          it.A = tc::operand(g.dY, 1, g.N); it.B = tc::operand(g.X, 1, g.K); it.out = g.dW; it.ldo = g.K; it.M = g.N; it.N = g.K; it.K = M;
          flops += 2.0 * g.N * g.K * M;
        } else {
          B200_TRY(linear_wgrad(g.dY, g.N, g.X, g.K, M, g.N, g.K, g.dW, st));
        }
      }
      if (bg[B_FC2_B]) cs.j[cs.count++] = ColsumJob{w.dyo[i + 1], bg[B_FC2_B], M, H};
      if (bg[B_FC1_B]) cs.j[cs.count++] = ColsumJob{w.dh[i], bg[B_FC1_B], M, F};
      if (bg[B_PROJ_B]) cs.j[cs.count++] = ColsumJob{w.dy1[i], bg[B_PROJ_B], M, H};
      if (bg[B_LN2_W] || bg[B_LN2_B]) ln.j[ln.count++] = LnParamJob{w.dln2[i], w.x1[i], w.ln2s[i], bg[B_LN2_W], bg[B_LN2_B]};
      if (bg[B_LN1_W] || bg[B_LN1_B]) ln.j[ln.count++] = LnParamJob{w.dln1[i], xin, w.ln1s[i], bg[B_LN1_W], bg[B_LN1_B]};
    }
    if constexpr (kTC) {
      if (ni) {
        B200_PROFD(st, "linear_wgrad grouped x%d mflop=%.0f", ni, flops * 1e-6);
        B200_TRY(tc::gemm_grouped(items, ni, st));
      }
    }
    B200_TRY(launch_colsum_multi<T>(cs, st));
    B200_TRY(launch_layernorm_bwd_params_multi<T>(ln, st));
    return 0;
  }

  // ------------------------------------------------------------ backward
  // G: gradient pointers indexed like P (null = not wanted).  Workspace must be the one the forward filled.
  int backward(const float* const* P, float* const* G, const float* x_in, char* ws, const float* d_enc4, const float* d_logits,
               int flags, cudaStream_t st) {
    defer_wg = n_grad_ev > 0 && !(flags & FLAG_INPLACE_WGRADS);
    deferred.clear();
    int rc = backward_impl(P, G, x_in, ws, d_enc4, d_logits, flags, st);
    if (defer_wg) {
      if (!rc) { B200_PROFC_BEGIN("B3 deferred conv wgrads", st); rc = flush_deferred(st); B200_PROFC_END(st); }
      deferred.clear();
      if (!rc) mark_grads(0, st);
    }
    return rc;
  }
  int backward_impl(const float* const* P, float* const* G, const float* x_in, char* ws, const float* d_enc4, const float* d_logits,
                    int flags, cudaStream_t st) {
    layout(ws, true);
    if (kTC) { B200_CHECK(packed_base, "bf16 mode needs the packed-weight buffer (b200_unetr_set_packed_weights)"); layout_packed(packed_base); }
    cur_params = P;
    int B = c.B, fs = c.fs;
    bool dec = (flags & FLAG_HAS_DLOGITS) && d_logits;
    bool enc = (flags & FLAG_NEED_ENCODER_GRAD) != 0;
    bool has_denc4 = (flags & FLAG_HAS_DENC4) && d_enc4;
    bool vit_from_top = false;  // does gradient reach blocks 10, 11 and the final LayerNorm?
    B200_CUDA(cudaMemsetAsync(w.dhs[0], 0, (size_t)((char*)(w.dhs[2] + (size_t)M * H) - (char*)w.dhs[0]), st));   // dhs[0..2] are consecutive
    B200_CUDA(cudaMemsetAsync(w.bwd_pool, 0, sizeof(double) * kBwdSlots * 3 * c.B * 8 * c.fs, st));
    B200_TRY(prezero_conv_grads(G, st));
    B200_PROFC_BEGIN("B1 head+decoders+encoders", st);

    if (dec) {
      int rows = (int)(B * V[0]);
      // head: d(d0) = dlogits^T W ; dW = dlogits d0 ; db = sum dlogits
      if (edge_co_ok(fs)) {
        B200_PROF("head_bwd", st);
        if (G[P_OUT_W]) B200_TRY(zero_grad(G[P_OUT_W], (size_t)c.ncls * fs, st));
        if (G[P_OUT_B]) B200_TRY(zero_grad(G[P_OUT_B], (size_t)c.ncls, st));
        long chunk = 2048;
        dim3 g((unsigned)((V[0] + chunk - 1) / chunk), B);
        size_t sm = head_bwd_smem(c.ncls, fs);
        static bool attr = false;
        if (!attr) {
          B200_CUDA(cudaFuncSetAttribute(head_bwd_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          B200_CUDA(cudaFuncSetAttribute(head_bwd_kernel<T, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          B200_CUDA(cudaFuncSetAttribute(head_bwd_kernel<T, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          B200_CUDA(cudaFuncSetAttribute(head_bwd2_kernel<T, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          attr = true;
        }
        if (c.ncls <= 16 && !getenv("B200_HEAD_BWD_V1")) {
          size_t sm2 = head_bwd2_smem(fs);
          if (fs == 8) head_bwd2_kernel<T, 8><<<g, 256, sm2, st>>>(d_logits, w.d0, P[P_OUT_W], c.ncls, V[0], chunk, w.gA, G[P_OUT_W], G[P_OUT_B]);
          else if (fs == 16) head_bwd2_kernel<T, 16><<<g, 256, sm2, st>>>(d_logits, w.d0, P[P_OUT_W], c.ncls, V[0], chunk, w.gA, G[P_OUT_W], G[P_OUT_B]);
          else head_bwd2_kernel<T, 32><<<g, 256, sm2, st>>>(d_logits, w.d0, P[P_OUT_W], c.ncls, V[0], chunk, w.gA, G[P_OUT_W], G[P_OUT_B]);
        }
        else if (fs == 8) head_bwd_kernel<T, 8><<<g, 256, sm, st>>>(d_logits, w.d0, P[P_OUT_W], c.ncls, V[0], chunk, w.gA, G[P_OUT_W], G[P_OUT_B]);
        else if (fs == 16) head_bwd_kernel<T, 16><<<g, 256, sm, st>>>(d_logits, w.d0, P[P_OUT_W], c.ncls, V[0], chunk, w.gA, G[P_OUT_W], G[P_OUT_B]);
        else head_bwd_kernel<T, 32><<<g, 256, sm, st>>>(d_logits, w.d0, P[P_OUT_W], c.ncls, V[0], chunk, w.gA, G[P_OUT_W], G[P_OUT_B]);
        B200_LAUNCH_CHECK();
      } else { B200_PROF("head_bwd", st);
      { RowIsOuter<NcdhwGather, true> al; al.g = {d_logits, c.ncls, V[0]};
        B200_TRY(launch_contract(al, ld2<float, true>(P[P_OUT_W], 1, fs), ep_plain<T>(w.gA, fs), rows, fs, c.ncls, 1, 1, st)); }
      if (G[P_OUT_W]) {
        B200_TRY(zero_grad(G[P_OUT_W], (size_t)c.ncls * fs, st));
        RowIsK<NcdhwGather, false> al; al.g = {d_logits, c.ncls, V[0]};
        EpAtomic ep = {G[P_OUT_W], (long)fs};
        B200_TRY(launch_contract(al, ld2<T, true>(w.d0, 1, fs), ep, c.ncls, fs, rows, 1, pick_splits(rows, 1), st));
      }
      if (G[P_OUT_B]) {
        B200_TRY(zero_grad(G[P_OUT_B], (size_t)c.ncls, st));
        rowsum_atomic_kernel<<<dim3(32, B * c.ncls), 256, 0, st>>>(d_logits, G[P_OUT_B], V[0], c.ncls);
        B200_LAUNCH_CHECK();
      }
      }
      // decoder2 block, encoder1 block, decoder2 transposed conv
      B200_TRY(res_bwd(cl<const T>(w.cat2, 2 * fs, 0, 2 * fs), 0, P[P_D2_C1], P[P_D2_C2], P[P_D2_C3], w.rs[4], cl<const T>(w.d0, fs, 0, fs),
                       cl<const T>(w.gA, fs, 0, fs), G[P_D2_C1], G[P_D2_C2], G[P_D2_C3], cl(w.dcat, 2 * fs, 0, 2 * fs), st));
      if (enc)
        B200_TRY(res_bwd(cl<const T>(w.xcl, c.Cin, 0, c.Cin), 0, P[P_E1_C1], P[P_E1_C2], P[P_E1_C3], w.rs[0], cl<const T>(w.cat2, 2 * fs, fs, fs),
                         cl<const T>(w.dcat, 2 * fs, fs, fs), G[P_E1_C1], G[P_E1_C2], G[P_E1_C3], cl<T>(nullptr, 0, 0, 0), st,
                         (edge_co_ok(fs) && G[P_E1_C1] && G[P_E1_C3]) ? x_in : nullptr));
      B200_TRY(convT_bwd(w.d1, 2 * fs, 2 * fs, 1, P[P_D2_T], cl<const T>(w.dcat, 2 * fs, 0, fs), G[P_D2_T], w.gA, 2 * fs, 0, st));
      // decoder3
      B200_TRY(res_bwd(cl<const T>(w.cat3, 4 * fs, 0, 4 * fs), 1, P[P_D3_C1], P[P_D3_C2], P[P_D3_C3], w.rs[3], cl<const T>(w.d1, 2 * fs, 0, 2 * fs),
                       cl<const T>(w.gA, 2 * fs, 0, 2 * fs), G[P_D3_C1], G[P_D3_C2], G[P_D3_C3], cl(w.dcat, 4 * fs, 0, 4 * fs), st));
      if (enc) {  // encoder2 chain: 48^3 <- 24^3 <- 12^3 <- tokens
        B200_TRY(convT_bwd(w.e2b, 2 * fs, 2 * fs, 2, P[P_E2_T2], cl<const T>(w.dcat, 4 * fs, 2 * fs, 2 * fs), G[P_E2_T2], w.dc2, 2 * fs, 0, st));
        B200_TRY(convT_bwd(w.e2a, 2 * fs, 2 * fs, 3, P[P_E2_T1], cl<const T>(w.dc2, 2 * fs, 0, 2 * fs), G[P_E2_T1], w.dc3, 2 * fs, 0, st));
        B200_TRY(convT_bwd(w.hsT[0], H, H, 4, P[P_E2_T0], cl<const T>(w.dc3, 2 * fs, 0, 2 * fs), G[P_E2_T0], w.dhs[0], H, 0, st));
      }
      B200_TRY(convT_bwd(w.d2, 4 * fs, 4 * fs, 2, P[P_D3_T], cl<const T>(w.dcat, 4 * fs, 0, 2 * fs), G[P_D3_T], w.gA, 4 * fs, 0, st));
      // decoder4
      B200_TRY(res_bwd(cl<const T>(w.cat4, 8 * fs, 0, 8 * fs), 2, P[P_D4_C1], P[P_D4_C2], P[P_D4_C3], w.rs[2], cl<const T>(w.d2, 4 * fs, 0, 4 * fs),
                       cl<const T>(w.gA, 4 * fs, 0, 4 * fs), G[P_D4_C1], G[P_D4_C2], G[P_D4_C3], cl(w.dcat, 8 * fs, 0, 8 * fs), st));
      if (enc) {
        B200_TRY(convT_bwd(w.e3a, 4 * fs, 4 * fs, 3, P[P_E3_T1], cl<const T>(w.dcat, 8 * fs, 4 * fs, 4 * fs), G[P_E3_T1], w.dc2, 4 * fs, 0, st));
        B200_TRY(convT_bwd(w.hsT[1], H, H, 4, P[P_E3_T0], cl<const T>(w.dc2, 4 * fs, 0, 4 * fs), G[P_E3_T0], w.dhs[1], H, 0, st));
      }
      B200_TRY(convT_bwd(w.d3, 8 * fs, 8 * fs, 3, P[P_D4_T], cl<const T>(w.dcat, 8 * fs, 0, 4 * fs), G[P_D4_T], w.gA, 8 * fs, 0, st));
      // decoder5
      B200_TRY(res_bwd(cl<const T>(w.cat5, 16 * fs, 0, 16 * fs), 3, P[P_D5_C1], P[P_D5_C2], P[P_D5_C3], w.rs[1], cl<const T>(w.d3, 8 * fs, 0, 8 * fs),
                       cl<const T>(w.gA, 8 * fs, 0, 8 * fs), G[P_D5_C1], G[P_D5_C2], G[P_D5_C3], cl(w.dcat, 16 * fs, 0, 16 * fs), st));
      // decoder5.transp_conv: weight grad always (it is a decoder parameter); input grad only when the encoder trains
      B200_TRY(convT_bwd(w.vit_out, H, H, 4, P[P_D5_T], cl<const T>(w.dcat, 16 * fs, 0, 8 * fs), G[P_D5_T], enc ? w.dvit : (T*)nullptr, H, 0, st));
      vit_from_top = enc;
    } else if (enc && has_denc4) {
      B200_CUDA(cudaMemsetAsync(w.dcat, 0, sizeof(T) * B * V[3] * 16 * fs, st));
    }
    if (!enc) return 0;
    // encoder4: gradient = skip half of d(concat5) (+ the external gradient of the returned enc4 tensor)
    if (has_denc4) B200_TRY(launch_layout<T>(d_enc4, nullptr, w.dcat, B, 8 * fs, V[3], 16 * fs, 8 * fs, 0, 1, st));
    if (dec || has_denc4)
      B200_TRY(convT_bwd(w.hsT[2], H, H, 4, P[P_E4_T0], cl<const T>(w.dcat, 16 * fs, 8 * fs, 8 * fs), G[P_E4_T0], w.dhs[2], H, 0, st));
    else
      return 0;

    // --- ViT backward
    if (!defer_wg) mark_grads(0, st);
    B200_PROFC_END(st); B200_PROFC_BEGIN("B2 vit+patch", st);
    float scale = 1.0f / sqrtf((float)dh);
    int top = 11;
    if (vit_from_top) {
      B200_TRY(launch_layernorm_bwd<T>(w.dvit, w.hs[11], w.lnfs, P[P_NORM_W], nullptr, w.dx, w.dyo[12], G[P_NORM_W], G[P_NORM_B], M, H, st));
    } else {
      top = 9;  // blocks 10, 11 and vit.norm are unreachable from enc4: their grads stay None
      B200_CUDA(cudaMemsetAsync(w.dx, 0, sizeof(float) * M * H, st));
      B200_CUDA(cudaMemsetAsync(w.dyo[10], 0, sizeof(T) * M * H, st));
    }
    // The loop below is the CRITICAL PATH only: dgrad GEMMs, attention backward and the LayerNorm input gradients.  Everything that
    // only feeds the optimizer (4 weight gradients, 3 bias sums and 2 LayerNorm parameter sums per block) reads per-block copies of
    // the output gradients and is issued by vit_param_grads() once per gradient group (blocks 8..11 / 4..7 / 0..3): 3 launches per
    // group instead of 36.
    int group_hi = top;
    const int gs = vit_group();
    for (int i = top; i >= 0; --i) {
      const float* const* bp = P + P_BLK0 + i * B_COUNT;
      const float* xin = i ? w.hs[i - 1] : w.x0;
      T* dyo = w.dyo[i + 1];      // d(hs[i]) as T (the fp32 original is w.dx)
      if (i == 9 || i == 6 || i == 3) {
        B200_TRY(launch_add_cast<T>(w.dx, w.dhs[(i - 3) / 3], dyo, (long)M * H, st));
      }
      // hs = x1 + fc2(h) + b2
      { EpStore<T> ep = ep_plain<T>(w.dh[i], F); ep.act = ACT_MUL_SAVED; ep.usrc = w.u[i];   // du = (dx W2) * gelu'(u), gelu'(u) saved by the forward epilogue
        B200_TRY(linear_dgrad<T>(dyo, H, bp[B_FC2_W], w.wfc2[i], M, H, F, ep, st)); }
      { SplitSum ss; memset(&ss, 0, sizeof(ss));
        if (can_split(M, H, F)) B200_TRY(linear_dgrad_split(w.dh[i], F, w.wfc1[i], M, F, H, &ss, st));
        else B200_TRY(linear_dgrad<T>(w.dh[i], F, bp[B_FC1_W], w.wfc1[i], M, F, H, ep_plain<T>(w.dln2[i], H), st));
        B200_TRY(launch_layernorm_bwd<T>(w.dln2[i], w.x1[i], w.ln2s[i], bp[B_LN2_W], w.dx, w.dx2, w.dy1[i], nullptr, nullptr, M, H, st, ss.nsplit ? &ss : nullptr)); }
      // x1 = xin + proj(att) + bp
      B200_TRY(linear_dgrad<T>(w.dy1[i], H, bp[B_PROJ_W], w.wproj[i], M, H, H, ep_plain<T>(w.datt, H), st));
      B200_TRY(attention_bwd(i, scale, st));
      { SplitSum ss; memset(&ss, 0, sizeof(ss));
        if (can_split(M, H, 3 * H)) B200_TRY(linear_dgrad_split(w.dqkv[i], 3 * H, w.wqkv[i], M, 3 * H, H, &ss, st));
        else B200_TRY(linear_dgrad<T>(w.dqkv[i], 3 * H, bp[B_QKV_W], w.wqkv[i], M, 3 * H, H, ep_plain<T>(w.dln1[i], H), st));
        B200_TRY(launch_layernorm_bwd<T>(w.dln1[i], xin, w.ln1s[i], bp[B_LN1_W], w.dx2, w.dx, w.dyo[i], nullptr, nullptr, M, H, st, ss.nsplit ? &ss : nullptr)); }
      if (i % gs == 0) {
        B200_TRY(vit_param_grads(G, i, group_hi, st));
        group_hi = i - 1;
        if (i) mark_grads((12 - i) / gs, st);      // the last group's event waits for the patch embedding below
      }
    }
    // --- patch embedding: dW = dx0^T rows(x), db = colsum, dpos = sum over batch
    if (G[P_PATCH_W]) {
      B200_PROF("patch_wgrad", st);
      int Kp = 4096 * c.Cin;
      if constexpr (kTC) {
        B200_TRY(tc::gemm(tc::operand(w.dyo[0], 1, H), tc::operand(w.apatch, 1, Kp), ep_plain<float>(G[P_PATCH_W], Kp), H, Kp, M, 1, 1, st));
      } else {
        RowIsK<PatchGather, true> bl; bl.g = {x_in, c.Cin, c.S0, c.S1, c.S2, g0, g1, g2, c.conv_patch};
        B200_TRY(launch_contract(ld2<float, true>(w.dx, 1, H), bl, ep_plain<float>(G[P_PATCH_W], Kp), H, Kp, M, 1, 1, st));
      }
    }
    if (G[P_PATCH_B]) B200_TRY(launch_colsum<float>(w.dx, G[P_PATCH_B], M, H, st));
    if (G[P_POS]) { batchsum_kernel<<<cdiv((long)L * H, 256), 256, 0, st>>>(w.dx, G[P_POS], B, (long)L * H); B200_LAUNCH_CHECK(); }
    mark_grads(12 / gs, st);
    B200_PROFC_END(st);
    return 0;
  }

  // debug: device pointer + element count of a named workspace buffer (valid after forward/backward laid it out)
  const void* peek(const char* name, size_t* bytes) const {
    struct { const char* n; const void* p; size_t b; } tab[] = {
      {"dc1", w.dc1, sizeof(T) * c.B * V[0] * c.fs}, {"da1", w.da1, sizeof(T) * c.B * V[0] * c.fs},
      {"dc2", w.dc2, sizeof(T) * c.B * V[0] * c.fs}, {"dc3", w.dc3, sizeof(T) * c.B * V[0] * c.fs},
      {"gA", w.gA, sizeof(T) * c.B * V[0] * c.fs}, {"dcat", w.dcat, sizeof(T) * c.B * V[0] * 2 * c.fs},
      {"a1_dec3", w.rs[3].a1, sizeof(T) * c.B * V[1] * 2 * c.fs}, {"c2_dec3", w.rs[3].c2, sizeof(T) * c.B * V[1] * 2 * c.fs},
      {"mr1_dec3", w.rs[3].mr1, sizeof(float) * 2 * c.B * 2 * c.fs}, {"cat3", w.cat3, sizeof(T) * c.B * V[1] * 4 * c.fs},
      {"bwd_acc", w.bwd_acc, sizeof(double) * 3 * c.B * 8 * c.fs}, {"d1", w.d1, sizeof(T) * c.B * V[1] * 2 * c.fs},
    };
    for (auto& e : tab) if (!strcmp(e.n, name)) { *bytes = e.b; return e.p; }
    return nullptr;
  }

  // transposed conv backward: dW (if non-null) and d(input) (if dx non-null) written as type TO rows [rows_in, Ci]
  template <class TO>
  int convT_bwd(const T* x, long ldx, int Ci, int in_level, const float* W, Cl<const T> dy, float* dW, TO* dx, long lddx, int accumulate, cudaStream_t st) {
    Sp s = sp(in_level);
    if constexpr (kTC) {
      const bf16* Wt = cur_params ? convT_packed(cur_params, W, true) : nullptr;
      constexpr int VN = Vec16<T>::N;
      if (Wt && dy.C % VN == 0 && dy.pitch % VN == 0 && dy.coff % VN == 0 && (ldx * 2) % 16 == 0) {
        int Co = dy.C, rows = (int)s.rows(), K8 = 8 * Co;
        { B200_PROFD(st, "convT_unshuffle %d @%d", Co, s.D);
          long tot = (long)rows * 8 * (Co / VN);
          const dim3 ug((unsigned)min(148L * 16, (tot + 255) / 256));
          if (tot + (long)ug.x * 256 < (1L << 32))
            B200_CUDA(launch_pdl(unshuffle_kernel<T, unsigned>, ug, dim3(256), 0, st, dy.p, ClView{dy.pitch, dy.coff}, Co, s.N, s.D, s.H, s.W, w.unsh));
          else
            B200_CUDA(launch_pdl(unshuffle_kernel<T, long>, ug, dim3(256), 0, st, dy.p, ClView{dy.pitch, dy.coff}, Co, s.N, s.D, s.H, s.W, w.unsh));
          B200_LAUNCH_CHECK(); }
        if (dW) {   // dW[ci][co*8+tap] = sum_v x[v,ci] U[v, tap*Co+co]   (voxels = reduction, split-K + atomics)
          B200_PROFD(st, "convT_wgrad %dx%d @%d", Ci, Co, s.D);
          B200_TRY(zero_grad(dW, (size_t)Ci * K8, st));
          EpAtomicTapRemap ep = {dW, Co};
          B200_TRY(tc::gemm(tc::operand(x, 1, ldx), tc::operand(w.unsh, 1, K8), ep, Ci, K8, rows, 1, 1, st, true));
        }
        if (dx) {   // dx[v,ci] = sum_j U[v,j] Wt[ci][j]
          B200_PROFD(st, "convT_dgrad %d<-%d @%d", Ci, Co, s.D);
          EpStore<TO> ep = ep_plain<TO>(dx, lddx); ep.accumulate = accumulate;
          B200_TRY(tc::gemm(tc::operand(w.unsh, K8, 1), tc::operand(Wt, K8, 1), ep, rows, Ci, K8, 1, 1, st));
        }
        return 0;
      }
    }
    if (dW) {
      B200_TRY(zero_grad(dW, (size_t)Ci * dy.C * 8, st));
      B200_TRY((simt_convT_wgrad<T, T>(x, ldx, Ci, dy, s, dW, st)));
    }
    if (dx) B200_TRY((simt_convT_dgrad<T, TO>(dy, s, W, Ci, dx, lddx, accumulate, st)));
    return 0;
  }
};

}  // namespace b200
