// extern "C" surface of libunetr_b200.so (declared in include/unetr_b200.h).
#include <math.h>
#include <stdarg.h>

#include <algorithm>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/unetr_b200.h"
#include "exec_iface.h"
#include "elementwise.cuh"
#include "edge_kernels.cuh"
#include "adamw.cuh"
#include "loss.cuh"
#include "sliding.cuh"
#include "metrics.cuh"
#include "augment.cuh"
#include "tc_gemm.cuh"
#include "tc_gemm_grouped.cuh"
#include "tc_attention.cuh"
#include "tc_conv_halo.cuh"
#include "tc_wgrad_halo.cuh"

namespace b200 {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

unsigned long long g_launches = 0;
bool g_prof_on = false;
bool g_prof_coarse = false;
long long* g_trace = nullptr;
int g_trace_n = 0, g_trace_cap = 0;
static std::vector<std::string> g_trace_tags;
void trace_tag(const char* fmt, ...) {
  char b[96]; va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof(b), fmt, ap); va_end(ap);
  g_trace_tags.push_back(b);
}
struct ProfRec { std::string tag; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
void prof_begin(const char* tag, cudaStream_t st) {
  ProfRec r; r.tag = tag;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) { cudaEventRecord(g_prof.back().b, st); }

static constexpr int P_COUNT_PUBLIC = 164;   // parameters in the C-ABI table (enum ParamIdx in exec.cuh; _lib.PARAM_COUNT)
struct Handle {
  UnetrConfig cfg;
  ExecIface* ex;
};
}  // namespace b200

using namespace b200;

extern "C" {

const char* b200_last_error(void) { return get_error(); }

int b200_device_check(void) {
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  B200_CUDA(cudaGetDeviceProperties(&p, dev));
  B200_CHECK(p.major == 10, "libunetr_b200 is built for sm_100a only; device %s is sm_%d%d", p.name, p.major, p.minor);
  return 0;
}

void* b200_unetr_create(const b200_unetr_config* c) {
  if (!c) { set_error("null config"); return nullptr; }
  if (c->img0 % 16 || c->img1 % 16 || c->img2 % 16 || c->img0 <= 0 || c->img1 <= 0 || c->img2 <= 0) {
    set_error("img_size must be positive multiples of the 16^3 patch (unetr.py:70)"); return nullptr;
  }
  if (c->hidden_size % c->num_heads) { set_error("hidden size should be divisible by num_heads."); return nullptr; }
  int fs = c->feature_size;
  if (fs < 8 || (fs & (fs - 1))) { set_error("feature_size must be a power of two >= 8 (got %d)", fs); return nullptr; }
  if (c->hidden_size % 8 || c->mlp_dim % 8) { set_error("hidden_size and mlp_dim must be multiples of 8"); return nullptr; }
  if (c->batch < 1 || c->in_channels < 1 || c->out_channels < 1 || c->out_channels > 32) {
    set_error("batch/in_channels/out_channels out of range (out_channels <= 32)"); return nullptr;
  }
  Handle* h = new (std::nothrow) Handle();
  if (!h) { set_error("out of host memory"); return nullptr; }
  h->cfg = UnetrConfig{c->batch, c->in_channels, c->out_channels, c->img0, c->img1, c->img2, fs, c->hidden_size, c->mlp_dim,
                       c->num_heads, c->conv_patch_embed, c->mode};
  h->ex = c->mode == 0 ? make_exec_f32(h->cfg) : make_exec_bf16(h->cfg);
  return h;
}
void b200_unetr_destroy(void* handle) {
  Handle* h = (Handle*)handle;
  if (!h) return;
  delete h->ex; delete h;
}
size_t b200_unetr_workspace_bytes(void* handle, int with_backward) {
  Handle* h = (Handle*)handle;
  return h->ex->workspace_bytes(with_backward != 0);
}
int b200_unetr_forward(void* handle, const float* const* params, const float* x, void* workspace, float* enc4_out,
                       float* logits_out, int flags, void* stream) {
  Handle* h = (Handle*)handle;
  B200_CHECK(h && params && x && workspace, "b200_unetr_forward: null argument");
  return h->ex->forward(params, x, (char*)workspace, enc4_out, logits_out, flags, (cudaStream_t)stream);
}
int b200_unetr_backward(void* handle, const float* const* params, float* const* grads, const float* x, void* workspace,
                        const float* d_enc4, const float* d_logits, int flags, void* stream) {
  Handle* h = (Handle*)handle;
  B200_CHECK(h && params && grads && x && workspace, "b200_unetr_backward: null argument");
  return h->ex->backward(params, grads, x, (char*)workspace, d_enc4, d_logits, flags, (cudaStream_t)stream);
}

void b200_unetr_set_grad_events(void* handle, void* const* events, int n) {
  Handle* h = (Handle*)handle;
  cudaEvent_t ev[13] = {};
  for (int i = 0; i < n && i < 13; ++i) ev[i] = (cudaEvent_t)events[i];
  h->ex->set_grad_events(ev, n);
}

size_t b200_unetr_packed_bytes(void* handle) { return ((Handle*)handle)->ex->packed_bytes(); }
void b200_unetr_set_packed_weights(void* handle, void* buf) { ((Handle*)handle)->ex->set_packed((char*)buf); }
int64_t b200_unetr_packed_cast_offset(void* handle, int param_index) {
  Handle* h = (Handle*)handle;
  if (!h || param_index < 0 || param_index >= P_COUNT_PUBLIC) return -1;
  return h->ex->packed_cast_offset(param_index);
}
int b200_unetr_pack_convs(void* handle, const float* const* params, void* packed, void* stream) {
  Handle* h = (Handle*)handle;
  B200_CHECK(h && params, "b200_unetr_pack_convs: null argument");
  return h->ex->pack_convs(params, (char*)packed, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- fused AdamW
/* tensors: device array of n_tensors {p, g, m, v, n, s0} (six 8-byte fields, adamw.cuh); chunks: device array of n_chunks {tensor, pad, start} covering every tensor in
 * pieces of b200_adamw_chunk() elements; step >= 1 is the 1-based update count (bias correction). */
long b200_adamw_chunk(void) { return kAdamChunk; }
int b200_adamw_step(const void* tensors, const void* chunks, int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay,
                    int step, void* stream) {
  return b200_adamw_step_capturable(tensors, chunks, n_chunks, lr, beta1, beta2, eps, weight_decay, step, nullptr, stream);
}
/* dev_step != NULL: the update count is the device int *dev_step, advanced by one before the update when step > 0 (step == 0: a
 * further launch of the same update, e.g. a later gradient range -- the count is read, not advanced): the launch sequence is then
 * identical every step and can be replayed from a CUDA graph. */
int b200_adamw_step_capturable(const void* tensors, const void* chunks, int n_chunks, float lr, float beta1, float beta2, float eps,
                               float weight_decay, int step, int* dev_step, void* stream) {
  B200_CHECK(tensors && chunks && n_chunks > 0 && (step >= 1 || dev_step), "b200_adamw_step: bad arguments");
  B200_PROF("adamw", (cudaStream_t)stream);
  const bool advance = dev_step && step > 0;     // step == 0 with dev_step: a further launch of the same update (a later gradient range)
  AdamHyper h;
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay;
  h.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  h.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  h.dev_step = dev_step;
  if (advance) { adam_step_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev_step); B200_LAUNCH_CHECK(); }
  adamw_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>((const AdamTensor*)tensors, (const AdamChunk*)chunks, h);
  B200_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- DiceCE
// staged kernels (loss.cuh): whole tiles of kDiceTile voxels, 16-byte aligned planes; B200_DICE_DIRECT=1 keeps the register kernels
static bool dice_staged_ok(const float* logits, const float* labels, const float* dlogits, int C, int64_t V) {
  static const bool off = getenv("B200_DICE_DIRECT") != nullptr;
  return !off && C <= 16 && V % kDiceTile == 0 && (((uintptr_t)logits | (uintptr_t)labels | (uintptr_t)dlogits) & 15) == 0;
}
static dim3 dice_staged_grid(int B, int64_t V) {
  const long tiles = V / kDiceTile;
  return dim3((unsigned)std::max(1L, std::min(tiles, (long)(2 * tc::num_sms() / B))), B);
}
static int dice_staged_attr() {
  static bool done = false;
  if (!done) {
    B200_CUDA(cudaFuncSetAttribute(dicece_staged_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dice_staged_smem(16)));
    B200_CUDA(cudaFuncSetAttribute(dicece_staged_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dice_staged_smem(16)));
    done = true;
  }
  return 0;
}
// scratch: double acc[B*C*3+1] | float coef[B*C*2]
size_t b200_dicece_scratch_bytes(int B, int C) { return sizeof(double) * ((size_t)B * C * 3 + 2) + sizeof(float) * (size_t)B * C * 2; }
int b200_dicece_forward(const float* logits, const float* labels, int B, int C, int64_t V, void* scratch, float* out3, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK(C >= 1 && C <= 32, "DiceCE supports 1..32 classes (got %d)", C);
  double* acc = (double*)scratch;
  float* coef = (float*)(acc + (size_t)B * C * 3 + 2);
  B200_PROF("dicece_fwd", st);
  B200_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * ((size_t)B * C * 3 + 2), st));
  dim3 g((unsigned)max(1L, min(148L * 8 / B + 1, (long)((V + 255) / 256))), B);
  // measured: the 4-voxel kernels need 184-254 registers and are ~15 % slower than the scalar ones at 14 classes; opt-in only
  const bool v4 = getenv("B200_DICE_V4") && V % 4 == 0 && C <= 16 && (((uintptr_t)logits | (uintptr_t)labels) & 15) == 0;
  dim3 g4((unsigned)max(1L, min(148L * 4 / B + 1, (long)((V / 4 + 255) / 256))), B);
  if (dice_staged_ok(logits, labels, nullptr, C, V)) {
    B200_TRY(dice_staged_attr());
    dicece_staged_kernel<16, false><<<dice_staged_grid(B, V), 256, dice_staged_smem(C), st>>>(logits, labels, C, V, acc, B * C, nullptr, nullptr, B, nullptr);
  }
  else if (v4) dicece_fwd4_kernel<16><<<g4, 256, 0, st>>>(logits, labels, C, V, acc, B * C);
  else if (C <= 16) dicece_fwd_kernel<16><<<g, 256, 0, st>>>(logits, labels, C, V, acc, B * C);
  else dicece_fwd_kernel<32><<<g, 256, 0, st>>>(logits, labels, C, V, acc, B * C);
  B200_LAUNCH_CHECK();
  dicece_finalize_kernel<<<1, 256, 0, st>>>(acc, B, C, V, out3, coef);
  B200_LAUNCH_CHECK();
  return 0;
}
int b200_dicece_backward(const float* logits, const float* labels, int B, int C, int64_t V, const void* scratch,
                         const float* upstream, float* dlogits, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK(C >= 1 && C <= 32, "DiceCE supports 1..32 classes (got %d)", C);
  const float* coef = (const float*)((const double*)scratch + (size_t)B * C * 3 + 2);
  B200_PROF("dicece_bwd", st);
  dim3 g((unsigned)max(1L, min(148L * 8 / B + 1, (long)((V + 255) / 256))), B);
  const bool v4 = getenv("B200_DICE_V4") && V % 4 == 0 && C <= 16 && (((uintptr_t)logits | (uintptr_t)labels | (uintptr_t)dlogits) & 15) == 0;
  dim3 g4((unsigned)max(1L, min(148L * 4 / B + 1, (long)((V / 4 + 255) / 256))), B);
  if (dice_staged_ok(logits, labels, dlogits, C, V)) {
    B200_TRY(dice_staged_attr());
    dicece_staged_kernel<16, true><<<dice_staged_grid(B, V), 256, dice_staged_smem(C), st>>>(logits, labels, C, V, nullptr, B * C, coef, upstream, B, dlogits);
  }
  else if (v4) dicece_bwd4_kernel<16><<<g4, 256, 0, st>>>(logits, labels, coef, upstream, B, C, V, dlogits);
  else if (C <= 16) dicece_bwd_kernel<16><<<g, 256, 0, st>>>(logits, labels, coef, upstream, B, C, V, dlogits);
  else dicece_bwd_kernel<32><<<g, 256, 0, st>>>(logits, labels, coef, upstream, B, C, V, dlogits);
  B200_LAUNCH_CHECK();
  return 0;
}

// DiceCELoss(to_onehot_y=False, sigmoid=True): target is the float multi-hot tensor [B][C][V] (seg:480); same scratch layout
int b200_dicece_sigmoid_forward(const float* logits, const float* target, int B, int C, int64_t V, void* scratch, float* out3, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK(C >= 1 && C <= 16, "sigmoid DiceCE supports 1..16 channels (got %d)", C);
  double* acc = (double*)scratch;
  float* coef = (float*)(acc + (size_t)B * C * 3 + 2);
  B200_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * ((size_t)B * C * 3 + 2), st));
  dim3 g((unsigned)max(1L, min(148L * 8 / B + 1, (long)((V + 255) / 256))), B);
  if (C <= 4) dicece_sig_fwd_kernel<4><<<g, 256, 0, st>>>(logits, target, C, V, acc, B * C);
  else dicece_sig_fwd_kernel<16><<<g, 256, 0, st>>>(logits, target, C, V, acc, B * C);
  B200_LAUNCH_CHECK();
  dicece_finalize_kernel<<<1, 256, 0, st>>>(acc, B, C, V, out3, coef);
  B200_LAUNCH_CHECK();
  return 0;
}
int b200_dicece_sigmoid_backward(const float* logits, const float* target, int B, int C, int64_t V, const void* scratch,
                                 const float* upstream, float* dlogits, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK(C >= 1 && C <= 16, "sigmoid DiceCE supports 1..16 channels (got %d)", C);
  const float* coef = (const float*)((const double*)scratch + (size_t)B * C * 3 + 2);
  dim3 g((unsigned)max(1L, min(148L * 8 / B + 1, (long)((V + 255) / 256))), B);
  if (C <= 4) dicece_sig_bwd_kernel<4><<<g, 256, 0, st>>>(logits, target, coef, upstream, B, C, V, dlogits);
  else dicece_sig_bwd_kernel<16><<<g, 256, 0, st>>>(logits, target, coef, upstream, B, C, V, dlogits);
  B200_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- validation metrics (DiceMetric / ConfusionMatrixMetric, seg:485-494)
int b200_seg_counts_onehot(const float* y_pred, const float* y, int B, int C, int64_t V, double* counts, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK(B >= 1 && C >= 1 && V >= 1, "b200_seg_counts_onehot: bad shape");
  B200_CUDA(cudaMemsetAsync(counts, 0, sizeof(double) * (size_t)B * C * 3, st));
  long rows = (long)B * C;
  dim3 g((unsigned)max(1L, min(148L * 8 / rows + 1, (long)((V / 4 + 255) / 256))), (unsigned)rows);
  seg_counts_onehot_kernel<<<g, 256, 0, st>>>(y_pred, y, V, counts);
  B200_LAUNCH_CHECK();
  return 0;
}
int b200_seg_counts_labels(const uint8_t* mask, const float* labels, int B, int C, int64_t V, double* counts, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK(B >= 1 && C >= 1 && C <= 32 && V >= 1, "b200_seg_counts_labels: 1..32 classes");
  B200_CUDA(cudaMemsetAsync(counts, 0, sizeof(double) * (size_t)B * C * 3, st));
  dim3 g((unsigned)max(1L, min(148L * 8 / B + 1, (long)((V + 1023) / 1024))), B);
  seg_counts_labels_kernel<<<g, 256, 0, st>>>(mask, labels, C, V, counts);
  B200_LAUNCH_CHECK();
  return 0;
}
int b200_seg_metrics(const double* counts, int N, int C, int64_t V, float* dice, float* confusion, void* stream) {
  B200_CHECK(N >= 1 && C >= 1, "b200_seg_metrics: bad shape");
  seg_metrics_kernel<<<cdiv((long)N * C, 256), 256, 0, (cudaStream_t)stream>>>(counts, N * C, (double)V, dice, confusion);
  B200_LAUNCH_CHECK();
  return 0;
}
int b200_metric_reduce(const float* f, int N, int C, int K, int reduction, float* out, float* not_nans, void* stream) {
  B200_CHECK(N >= 1 && C >= 1 && K >= 1 && (reduction == 0 || reduction == 1), "b200_metric_reduce: reduction 0 (mean) or 1 (mean_batch)");
  metric_reduce_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(f, N, C, K, reduction, out, not_nans);
  B200_LAUNCH_CHECK();
  return 0;
}
int b200_confusion_metric(const float* cm, int rows, int metric, float* out, void* stream) {
  B200_CHECK(rows >= 1 && (metric == 0 || metric == 1), "b200_confusion_metric: metric 0 (precision) or 1 (sensitivity)");
  confusion_metric_kernel<<<cdiv(rows, 256), 256, 0, (cudaStream_t)stream>>>(cm, rows, metric, out);
  B200_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- ranking loss
// scratch: double gram[C*256] | double loss | float coef[C*256]
size_t b200_ranking_scratch_bytes(int C) { return sizeof(double) * ((size_t)C * 256 + 2) + sizeof(float) * (size_t)C * 256; }
static RankGeom to_geom(const b200_rank_geom* g) {
  RankGeom r;
  for (int i = 0; i < 4; ++i) { r.src[i] = g->src[i]; r.grad[i] = g->grad[i]; r.idx[i] = g->idx[i]; }
  r.sc = g->stride_c; r.ss = g->stride_slice; r.sf0 = g->stride_f0; r.sf1 = g->stride_f1;
  r.C = g->channels; r.F0 = g->f0; r.F1 = g->f1; r.temperature = g->temperature;
  return r;
}
int b200_ranking_forward(const b200_rank_geom* g, void* scratch, float* loss_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK(g && scratch && loss_out, "b200_ranking_forward: null argument");
  RankGeom r = to_geom(g);
  double* gram = (double*)scratch; double* loss = gram + (size_t)r.C * 256;
  float* coef = (float*)(loss + 2);
  B200_CUDA(cudaMemsetAsync(gram, 0, sizeof(double) * ((size_t)r.C * 256 + 2), st));
  int F = r.F0 * r.F1;
  int splits = max(1, min(cdiv(F, 64), cdiv(148 * 2, r.C)));
  rank_gram_kernel<<<dim3(r.C, splits), 256, 0, st>>>(r, gram);
  B200_LAUNCH_CHECK();
  rank_loss_kernel<<<r.C, 576, 0, st>>>(gram, r.C, r.temperature, loss, coef);
  B200_LAUNCH_CHECK();
  cast_kernel<double, float><<<1, 32, 0, st>>>(loss, loss_out, 1);
  B200_LAUNCH_CHECK();
  return 0;
}
int b200_ranking_backward(const b200_rank_geom* g, const void* scratch, const float* upstream, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  RankGeom r = to_geom(g);
  const float* coef = (const float*)((const double*)scratch + (size_t)r.C * 256 + 2);
  int F = r.F0 * r.F1;
  rank_grad_kernel<<<dim3(r.C, cdiv(F, 256)), 256, 0, st>>>(r, coef, upstream);
  B200_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- sliding window
static SwGeom to_sw(const b200_sw_geom* g) {
  return SwGeom{g->channels, g->d, g->h, g->w, g->pad_d, g->pad_h, g->pad_w, g->padded_d, g->padded_h, g->padded_w, g->roi0, g->roi1, g->roi2};
}
int b200_sw_gather(const float* volume, float* windows, const b200_sw_geom* g, const int32_t* starts, int n, float cval, void* stream) {
  B200_CHECK(n >= 1 && n <= 16, "sw_gather takes 1..16 windows per call");
  SwBatch wb; wb.n = n;
  for (int i = 0; i < n; ++i) wb.w[i] = SwWindow{starts[4 * i], starts[4 * i + 1], starts[4 * i + 2], starts[4 * i + 3]};
  SwGeom sg = to_sw(g);
  long total = (long)sg.C * sg.r0 * sg.r1 * sg.r2 * n;
  sw_gather_kernel<<<(unsigned)min(148L * 16, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(volume, windows, sg, wb, cval);
  B200_LAUNCH_CHECK();
  return 0;
}
int b200_sw_accumulate(float* acc, const float* pred, const b200_sw_geom* g, const int32_t* s, void* stream) {
  SwGeom sg = to_sw(g);
  long per = (long)sg.C * sg.r0 * sg.r1 * sg.r2;
  sw_accumulate_kernel<<<(unsigned)min(148L * 16, (per + 255) / 256), 256, 0, (cudaStream_t)stream>>>(acc, pred, sg, SwWindow{s[0], s[1], s[2], s[3]});
  B200_LAUNCH_CHECK();
  return 0;
}
int b200_sw_accumulate_n(float* acc, const float* pred, const b200_sw_geom* g, const int32_t* starts, int n, void* stream) {
  B200_CHECK(n >= 1 && n <= 16, "sw_accumulate_n takes 1..16 windows per call");
  SwGeom sg = to_sw(g);
  SwBatch wb; wb.n = n;
  SwBox bx; int x1 = 0, y1 = 0, z1 = 0;
  for (int i = 0; i < n; ++i) {
    wb.w[i] = SwWindow{starts[4 * i], starts[4 * i + 1], starts[4 * i + 2], starts[4 * i + 3]};
    B200_CHECK(wb.w[i].b == wb.w[0].b, "sw_accumulate_n: the windows of one call belong to one batch item");
    const int a = wb.w[i].s0, b = wb.w[i].s1, c = wb.w[i].s2;
    if (i == 0) { bx.x0 = a; bx.y0 = b; bx.z0 = c; x1 = a; y1 = b; z1 = c; }
    bx.x0 = std::min(bx.x0, a); bx.y0 = std::min(bx.y0, b); bx.z0 = std::min(bx.z0, c);
    x1 = std::max(x1, a); y1 = std::max(y1, b); z1 = std::max(z1, c);
  }
  bx.nx = x1 - bx.x0 + sg.r0; bx.ny = y1 - bx.y0 + sg.r1; bx.nz = z1 - bx.z0 + sg.r2;
  bool v4 = sg.r2 % 4 == 0 && sg.PW % 4 == 0 && bx.z0 % 4 == 0 && bx.nz % 4 == 0 && (((uintptr_t)acc | (uintptr_t)pred) & 15) == 0;
  for (int i = 0; i < n; ++i) v4 = v4 && wb.w[i].s2 % 4 == 0;
  B200_CHECK(sg.C <= 65535, "sw_accumulate_n: at most 65535 channels");
  const dim3 grid((unsigned)bx.nx, (unsigned)sg.C);      // one block per (row, channel) plane of the bounding box
  if (v4 && n <= 4) sw_accumulate_multi_kernel<4, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(acc, pred, sg, wb, bx);
  else if (v4) sw_accumulate_multi_kernel<4, 16><<<grid, 256, 0, (cudaStream_t)stream>>>(acc, pred, sg, wb, bx);
  else sw_accumulate_multi_kernel<1, 16><<<grid, 256, 0, (cudaStream_t)stream>>>(acc, pred, sg, wb, bx);
  B200_LAUNCH_CHECK();
  return 0;
}
static int sw_finalize_impl(const float* acc, float* out, uint8_t* mask, const b200_sw_geom* g, int batch, const int32_t* s0, int n0,
                            const int32_t* s1, int n1, const int32_t* s2, int n2, const float* labels, double* counts, const SwSlab* slab,
                            int zero_counts, void* stream);
int b200_sw_finalize_metric(const float* acc, float* out, uint8_t* mask, const b200_sw_geom* g, int batch, const int32_t* s0, int n0,
                            const int32_t* s1, int n1, const int32_t* s2, int n2, const float* labels, double* counts, void* stream) {
  return sw_finalize_impl(acc, out, mask, g, batch, s0, n0, s1, n1, s2, n2, labels, counts, nullptr, 1, stream);
}
/* slab form (multi-GPU): this rank owns un-padded rows [d0, d0 + nd) of ONE batch item (index `item` of the full-size mask / labels /
 * counts); acc holds padded rows [acc_xoff, acc_xoff + acc_rows) of that item; out (nullable) holds un-padded rows
 * [out_d0, out_d0 + out_rows).  counts (if given) are accumulated, not zeroed: the caller zeroes them once and all-reduces. */
int b200_sw_finalize_slab(const float* acc, float* out, uint8_t* mask, const b200_sw_geom* g, int item, const int32_t* s0, int n0,
                          const int32_t* s1, int n1, const int32_t* s2, int n2, const float* labels, double* counts,
                          int d0, int nd, int acc_xoff, int acc_rows, int out_d0, int out_rows, void* stream) {
  B200_CHECK(item >= 0 && nd >= 0 && d0 >= 0 && d0 + nd <= g->d, "b200_sw_finalize_slab: rows out of range");
  if (nd == 0) return 0;
  SwSlab sl = {d0, nd, acc_xoff, acc_rows, out_d0, out_rows};
  const long vox = (long)g->d * g->h * g->w;
  return sw_finalize_impl(acc, out, mask ? mask + (size_t)item * vox : nullptr, g, 1, s0, n0, s1, n1, s2, n2,
                          labels ? labels + (size_t)item * vox : nullptr, counts ? counts + (size_t)item * g->channels * 3 : nullptr, &sl, 0, stream);
}
static int sw_finalize_impl(const float* acc, float* out, uint8_t* mask, const b200_sw_geom* g, int batch, const int32_t* s0, int n0,
                            const int32_t* s1, int n1, const int32_t* s2, int n2, const float* labels, double* counts, const SwSlab* slab,
                            int zero_counts, void* stream) {
  B200_CHECK(n0 <= 64 && n1 <= 64 && n2 <= 64, "more than 64 window starts along one axis");
  B200_CHECK((labels == nullptr) == (counts == nullptr), "labels and counts go together");
  SwStarts st; st.n0 = n0; st.n1 = n1; st.n2 = n2;
  for (int i = 0; i < n0; ++i) st.s0[i] = s0[i];
  for (int i = 0; i < n1; ++i) st.s1[i] = s1[i];
  for (int i = 0; i < n2; ++i) st.s2[i] = s2[i];
  SwGeom sg = to_sw(g);
  B200_CHECK(!counts || sg.C <= 32, "fused validation counts take at most 32 classes");
  if (counts && zero_counts) B200_CUDA(cudaMemsetAsync(counts, 0, sizeof(double) * (size_t)batch * sg.C * 3, (cudaStream_t)stream));
  SwSlab sl = slab ? *slab : SwSlab{0, sg.D, 0, sg.PD, 0, sg.D};
  long vox = (long)sl.nd * sg.H * sg.W;
  dim3 grid((unsigned)max(1L, min(148L * 16 / batch + 1, (vox + 255) / 256)), batch);
  sw_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(acc, out, mask, sg, st, labels, counts, sl);
  B200_LAUNCH_CHECK();
  return 0;
}
/* pieces: n x {s0, s1, s2, x_lo, x_hi, nx, xbase} (int32) + preds: n device pointers; acc = slab [C][nrows][PH][PW] of padded rows
 * [xoff, xoff + nrows).  Adds the pieces in the order given (= global window order). */
int b200_sw_accumulate_slab(float* acc, const b200_sw_geom* g, const void* const* preds, const int32_t* pieces, int n, int xoff, int nrows,
                            void* stream) {
  B200_CHECK(n >= 1 && n <= 16, "sw_accumulate_slab takes 1..16 pieces per call");
  SwGeom sg = to_sw(g);
  SwPieces ps; ps.n = n;
  SwBox bx; int x1 = 0, y1 = 0, z1 = 0;
  for (int i = 0; i < n; ++i) {
    const int32_t* q = pieces + 7 * i;
    ps.p[i] = SwPiece{(const float*)preds[i], q[0], q[1], q[2], q[3], q[4], q[5], q[6]};
    B200_CHECK(q[3] >= q[0] && q[4] <= q[0] + sg.r0 && q[6] <= q[3] && q[6] + q[5] >= q[4], "sw_accumulate_slab: piece %d rows outside its window/buffer", i);
    const int a = q[3], a1 = q[4], b = q[1], c = q[2];
    if (i == 0) { bx.x0 = a; x1 = a1; bx.y0 = b; y1 = b; bx.z0 = c; z1 = c; }
    bx.x0 = std::min(bx.x0, a); x1 = std::max(x1, a1);
    bx.y0 = std::min(bx.y0, b); y1 = std::max(y1, b); bx.z0 = std::min(bx.z0, c); z1 = std::max(z1, c);
  }
  bx.x0 = std::max(bx.x0, xoff); x1 = std::min(x1, xoff + nrows);
  bx.nx = x1 - bx.x0; bx.ny = y1 - bx.y0 + sg.r1; bx.nz = z1 - bx.z0 + sg.r2;
  if (bx.nx <= 0) return 0;
  bool v4 = sg.r2 % 4 == 0 && sg.PW % 4 == 0 && bx.z0 % 4 == 0 && bx.nz % 4 == 0 && ((uintptr_t)acc & 15) == 0;
  for (int i = 0; i < n; ++i) v4 = v4 && ps.p[i].s2 % 4 == 0 && ((uintptr_t)ps.p[i].pred & 15) == 0;
  B200_CHECK(sg.C <= 65535, "sw_accumulate_slab: at most 65535 channels");
  const dim3 grid((unsigned)bx.nx, (unsigned)sg.C);      // one block per (row, channel) plane of the bounding box
  if (v4 && n <= 4) sw_accumulate_slab_kernel<4, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(acc, sg, ps, bx, xoff, nrows);
  else if (v4) sw_accumulate_slab_kernel<4, 16><<<grid, 256, 0, (cudaStream_t)stream>>>(acc, sg, ps, bx, xoff, nrows);
  else sw_accumulate_slab_kernel<1, 16><<<grid, 256, 0, (cudaStream_t)stream>>>(acc, sg, ps, bx, xoff, nrows);
  B200_LAUNCH_CHECK();
  return 0;
}
/* rows [x_from, x_from + nx) of a window prediction [C][r0][r1][r2] -> contiguous [C][nx][r1][r2] (the wire format of a halo piece):
 * one strided device-to-device copy on the DMA engine */
int b200_sw_pack_rows(const float* pred, float* dst, const b200_sw_geom* g, int x_from, int nx, void* stream) {
  B200_CHECK(x_from >= 0 && nx >= 1 && x_from + nx <= g->roi0, "b200_sw_pack_rows: rows outside the window");
  const size_t plane = (size_t)g->roi1 * g->roi2 * sizeof(float);
  B200_CUDA(cudaMemcpy2DAsync(dst, plane * nx, pred + (size_t)x_from * g->roi1 * g->roi2, plane * g->roi0, plane * nx, g->channels,
                              cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
int b200_sw_finalize(const float* acc, float* out, uint8_t* mask, const b200_sw_geom* g, int batch, const int32_t* s0, int n0,
                     const int32_t* s1, int n1, const int32_t* s2, int n2, void* stream) {
  return b200_sw_finalize_metric(acc, out, mask, g, batch, s0, n0, s1, n1, s2, n2, nullptr, nullptr, stream);
}

// ---------------------------------------------------------------- GPU-side crop sampling / augmentation
int b200_aug_blocks(int64_t voxels) { return (int)((voxels + kAugBlock - 1) / kAugBlock); }
/* prefix: int32[2][b200_aug_blocks(V)] (out: exclusive prefix of the per-block foreground / background counts); totals: int64[2] (device) */
int b200_aug_index(const float* label, int label_channels, const float* image, int image_channels, float image_threshold, int64_t voxels,
                   int32_t* prefix, int64_t* totals, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK(label && label_channels >= 1 && voxels >= 1 && voxels < (1LL << 31) && prefix && totals, "b200_aug_index: bad arguments");
  const int nblk = b200_aug_blocks(voxels);
  aug_fgbg_counts_kernel<<<nblk, 256, 0, st>>>(label, label_channels, image, image ? image_channels : 0, image_threshold, voxels, nblk, prefix);
  B200_LAUNCH_CHECK();
  aug_prefix_kernel<<<2, 1024, 0, st>>>(prefix, nblk, (long long*)totals);
  B200_LAUNCH_CHECK();
  return 0;
}
/* picks: n x {use_fg, k} (int64, host); starts: device int32[n][3] receives the crop start of each pick (RandCropByPosNegLabeld) */
int b200_aug_pick_centers(const float* label, int label_channels, const float* image, int image_channels, float image_threshold, int d, int h,
                          int w, const int32_t* prefix, const int64_t* picks, int n, int roi0, int roi1, int roi2, int32_t* starts, void* stream) {
  B200_CHECK(n >= 1 && n <= 16, "b200_aug_pick_centers takes 1..16 picks per call");
  B200_CHECK(roi0 <= d && roi1 <= h && roi2 <= w, "The size of the proposed random crop ROI is larger than the image size.");
  AugPicks pk; pk.n = n;
  for (int i = 0; i < n; ++i) { pk.p[i].use_fg = (int)picks[2 * i]; pk.p[i].k = picks[2 * i + 1]; }
  AugDims g = {d, h, w, roi0, roi1, roi2};
  aug_pick_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(label, label_channels, image, image ? image_channels : 0, image_threshold, prefix,
                                                        b200_aug_blocks((int64_t)d * h * w), pk, g, starts);
  B200_LAUNCH_CHECK();
  return 0;
}
/* starts: device int32[n][3]; maps: n x {perm[3], flip[3], shift} (host, 7 x 4 bytes each); out_image / out_label may be NULL */
int b200_aug_crop(const float* image, int image_channels, const float* label, int label_channels, int d, int h, int w, const int32_t* starts,
                  const b200_aug_map* maps, int n, int roi0, int roi1, int roi2, int brats, float* out_image, float* out_label, void* stream) {
  B200_CHECK(n >= 1 && n <= 16, "b200_aug_crop takes 1..16 samples per call");
  B200_CHECK(!brats || label_channels == 1, "the BraTS conversion takes a single-channel label map");
  AugMaps mp; mp.n = n;
  for (int i = 0; i < n; ++i) {
    int seen = 0;
    for (int a = 0; a < 3; ++a) { mp.m[i].perm[a] = maps[i].perm[a]; mp.m[i].flip[a] = maps[i].flip[a]; if (maps[i].perm[a] >= 0 && maps[i].perm[a] < 3) seen |= 1 << maps[i].perm[a]; }
    B200_CHECK(seen == 7, "b200_aug_crop: perm of sample %d is not a permutation of the three axes", i);
    const int ext[3] = {roi0, roi1, roi2};
    for (int a = 0; a < 3; ++a) B200_CHECK(ext[a] == ext[maps[i].perm[a]], "b200_aug_crop: rot90 of a non-square crop plane changes the crop shape");
    mp.m[i].shift = maps[i].shift;
  }
  AugDims g = {d, h, w, roi0, roi1, roi2};
  const long per = (long)roi0 * roi1 * roi2;
  dim3 grid((unsigned)std::max(1L, std::min(148L * 8 / n + 1, (per + 255) / 256)), n);
  aug_crop_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(image, image_channels, label, label_channels, g, starts, mp, brats, out_image, out_label);
  B200_LAUNCH_CHECK();
  return 0;
}

int b200_unetr_peek(void* handle, const char* name, void* dst, size_t cap) {
  Handle* h = (Handle*)handle;
  size_t bytes = 0;
  const void* src = h->ex->peek(name, &bytes);
  B200_CHECK(src, "no workspace buffer named %s", name);
  if (bytes > cap) bytes = cap;
  B200_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToDevice));
  return 0;
}
unsigned long long b200_launch_count(void) { return g_launches; }
void b200_prof_enable(int on) { g_prof_on = on == 1; g_prof_coarse = on == 2; }
/* synchronises the device, writes "tag ms count\n" lines (sorted by time) and clears the records */
int b200_prof_report(char* buf, int cap) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<double, int>> acc;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
    auto& e = acc[r.tag]; e.first += ms; e.second += 1;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  std::vector<std::pair<double, std::string>> v;
  for (auto& kv : acc) v.push_back({kv.second.first, kv.first});
  std::sort(v.rbegin(), v.rend());
  int off = 0;
  for (auto& e : v) {
    int n = snprintf(buf + off, cap - off, "%s %.4f %d\n", e.second.c_str(), e.first, acc[e.second].second);
    if (n < 0 || off + n >= cap) break;
    off += n;
  }
  if (cap > 0) buf[off < cap ? off : cap - 1] = 0;
  return 0;
}

/* x: bf16 channels-last [N,D,H,W,in_pitch] window (in_coff, Ci); w: fp32 [Co][Ci][ks^3]; out bf16 [N,D,H,W,out_pitch];
 * dgrad!=0 runs the transposed/flipped packing (input has Co channels, output Ci); stats: double[N*Cout*2] or null;
 * scratch: 2*Co*Ci*ks^3 bf16 */
int b200_test_tc_conv(const void* x, int in_pitch, int in_coff, int Ci, int N, int D, int H, int W, const float* w, int Co, int ks,
                      void* out, int out_pitch, int out_coff, int accumulate, int dgrad, double* stats, void* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int taps = ks * ks * ks;
  bf16* wf = (bf16*)scratch; bf16* wd = wf + (size_t)Co * Ci * taps;
  pack_conv_weights_kernel<<<64, 256, 0, st>>>(w, wf, wd, Co, Ci, taps);
  B200_LAUNCH_CHECK();
  if (!dgrad) {
    B200_CHECK(tc::conv_supported(Ci, Co, in_pitch, in_coff, out_pitch, out_coff), "shape unsupported by the tcgen05 conv");
    if (ks == 3 && !getenv("B200_TEST_NO_HALO") && tc::conv_halo_supported(Ci, Co))
      return tc::conv_halo((const bf16*)x, in_pitch, in_coff, Ci, N, D, H, W, wf, Co, (bf16*)out, out_pitch, out_coff, accumulate, stats, st);
    return tc::conv((const bf16*)x, in_pitch, in_coff, Ci, N, D, H, W, wf, Co, ks, (bf16*)out, out_pitch, out_coff, accumulate, stats, st);
  }
  B200_CHECK(tc::conv_supported(Co, Ci, in_pitch, in_coff, out_pitch, out_coff), "shape unsupported by the tcgen05 conv");
  if (ks == 3 && !getenv("B200_TEST_NO_HALO") && tc::conv_halo_supported(Co, Ci))
    return tc::conv_halo((const bf16*)x, in_pitch, in_coff, Co, N, D, H, W, wd, Ci, (bf16*)out, out_pitch, out_coff, accumulate, stats, st);
  return tc::conv((const bf16*)x, in_pitch, in_coff, Co, N, D, H, W, wd, Ci, ks, (bf16*)out, out_pitch, out_coff, accumulate, stats, st);
}

/* dgrad of a 3^3 convolution (dy: [N,D,H,W,Co] bf16 dense, w fp32 [Co][Ci][27] -> out = dx [N,D,H,W,Ci] bf16) with the first pass of the
 * InstanceNorm + LeakyReLU backward folded into the epilogue: act = lrelu(norm(.)) saved by the forward ([N,D,H,W,Ci] bf16), acc double
 * [N][Ci][3] receives (sum g, sum g*n, untouched), g = dx * lrelu'(act), n recovered from act.  Returns 0 and *folded = 1 when the kernel
 * that ran supports the fold (tc_conv_halo48.cuh), *folded = 0 otherwise (acc untouched).  scratch: 2*Co*Ci*27 bf16 */
int b200_test_tc_conv_dgrad_normbwd(const void* dy, int Co, int Ci, int N, int D, int H, int W, const float* w, const void* act, void* out, double* acc,
                                    int* folded, void* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  bf16* wf = (bf16*)scratch; bf16* wd = wf + (size_t)Co * Ci * 27;
  pack_conv_weights_kernel<<<64, 256, 0, st>>>(w, wf, wd, Co, Ci, 27);
  B200_LAUNCH_CHECK();
  B200_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 3 * N * Ci, st));
  bool done = false;
  tc::HaloNormBwd nb = {(const bf16*)act, Ci, 0, acc, &done};
  B200_CHECK(tc::conv_halo_supported(Co, Ci), "shape unsupported by the halo conv");
  B200_TRY(tc::conv_halo((const bf16*)dy, Co, 0, Co, N, D, H, W, wd, Ci, (bf16*)out, Ci, 0, 0, nullptr, st, nullptr, 0, &nb));
  *folded = done ? 1 : 0;
  return 0;
}

/* Fused 3^3 + 1^3 convolution of the residual block (tc_conv_halo.cuh, HaloParams::mode2), dense channels-last bf16 tensors:
 *  mode 1 (forward):  out = conv3(x; w3), out2 = conv1(x; w1), stats / stats2 = per-(n,c) sum and sum of squares   (x has Ci channels)
 *  mode 2 (dgrad):    out = dgrad3(x; w3) + dgrad1(x2; w1)                                  (x, x2 have Co channels, out has Ci)
 * w3 fp32 [Co][Ci][27], w1 fp32 [Co][Ci]; scratch: 2*Co*Ci*28 bf16 */
int b200_test_tc_conv_fused(const void* x, const void* x2, int Ci, int Co, int N, int D, int H, int W, const float* w3, const float* w1, int mode,
                            void* out, void* out2, double* stats, double* stats2, void* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  bf16* w3f = (bf16*)scratch; bf16* w3d = w3f + (size_t)Co * Ci * 27; bf16* w1f = w3d + (size_t)Co * Ci * 27; bf16* w1d = w1f + (size_t)Co * Ci;
  pack_conv_weights_kernel<<<64, 256, 0, st>>>(w3, w3f, w3d, Co, Ci, 27);
  B200_LAUNCH_CHECK();
  pack_conv_weights_kernel<<<64, 256, 0, st>>>(w1, w1f, w1d, Co, Ci, 1);
  B200_LAUNCH_CHECK();
  if (mode == 1) {
    B200_CHECK(tc::conv_halo_fused_supported(Ci, Co, 1), "fused forward unsupported for %d->%d", Ci, Co);
    tc::HaloFused fu = {1, w1f, (bf16*)out2, Co, 0, stats2, nullptr, 0, 0};
    return tc::conv_halo((const bf16*)x, Ci, 0, Ci, N, D, H, W, w3f, Co, (bf16*)out, Co, 0, 0, stats, st, &fu);
  }
  B200_CHECK(mode == 2 && tc::conv_halo_fused_supported(Co, Ci, 2), "fused dgrad unsupported for %d<-%d", Ci, Co);
  tc::HaloFused fu = {2, w1d, nullptr, 0, 0, nullptr, (const bf16*)x2, Co, 0};
  return tc::conv_halo((const bf16*)x, Co, 0, Co, N, D, H, W, w3d, Ci, (bf16*)out, Ci, 0, 0, nullptr, st, &fu);
}

/* dW fp32 [Co][Ci][ks^3] = sum_v dy[v,co] x[v+tap,ci]; x, dy channels-last bf16 windows */
int b200_test_tc_wgrad(const void* x, int x_pitch, int x_coff, int Ci, const void* dy, int dy_pitch, int dy_coff, int Co, int N, int D, int H,
                       int W, int ks, float* dW, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK(tc::wgrad_supported(Ci, Co, x_pitch, x_coff, dy_pitch, dy_coff), "shape unsupported by the tcgen05 wgrad");
  B200_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * Co * Ci * ks * ks * ks, st));
  if (tc::wgrad_halo_supported(Ci, Co, ks))
    return tc::conv_wgrad_halo((const bf16*)x, x_pitch, x_coff, Ci, (const bf16*)dy, dy_pitch, dy_coff, Co, N, D, H, W, dW, st);
  // the deterministic epilogue needs a scratch buffer for the partial tiles: the hook keeps one (grown on demand, never freed)
  static float* scratch = nullptr; static size_t scratch_bytes = 0;
  const size_t need = tc::wgrad_scratch_bytes(Ci, Co, ks, N, D, H, W);
  if (need > scratch_bytes) {
    B200_CUDA(cudaStreamSynchronize(st));
    if (scratch) B200_CUDA(cudaFree(scratch));
    B200_CUDA(cudaMalloc(&scratch, need)); scratch_bytes = need;
  }
  return tc::conv_wgrad((const bf16*)x, x_pitch, x_coff, Ci, (const bf16*)dy, dy_pitch, dy_coff, Co, N, D, H, W, ks, dW, st, scratch, scratch_bytes);
}

/* in-situ trace of the tcgen05 launches: buf = device int64[2*cap] pre-filled with (INT64_MAX, 0) pairs; null stops tracing */
void b200_trace_begin(void* buf, int cap) { g_trace = (long long*)buf; g_trace_cap = cap; g_trace_n = 0; g_trace_tags.clear(); }
int b200_trace_count(void) { return g_trace_n; }
int b200_trace_tags(char* out, int cap) {
  int off = 0;
  for (auto& t : g_trace_tags) { int n = snprintf(out + off, cap - off, "%s\n", t.c_str()); if (n < 0 || off + n >= cap) break; off += n; }
  if (cap > 0) out[off < cap ? off : cap - 1] = 0;
  return 0;
}

// ---------------------------------------------------------------- op-level hooks for the backward element-wise kernels
// (each kernel of the backward pinned on its own against fp32 torch with an injected upstream gradient: tests/test_gpu_ops_bwd.py)
}  // extern "C"
template <class T>
static int test_layernorm_bwd_t(const void* g, const float* x, const float* stats, const float* gamma, const float* dx_res, float* dx_out,
                                void* dx_cast, float* dgamma, float* dbeta, int M, int H, cudaStream_t st) {
  return launch_layernorm_bwd<T>((const T*)g, x, stats, gamma, dx_res, dx_out, (T*)dx_cast, dgamma, dbeta, M, H, st);
}
extern "C" {
/* LayerNorm backward (exec.cuh transformer blocks): g [M,H] (bf16 when bf16 != 0, else fp32), x fp32 [M,H], stats fp32 [M][2] = (mean,
 * rstd), gamma [H]; dx_out fp32 = dx_res (nullable) + d/dx; dx_cast = the same as T; dgamma / dbeta [H] */
int b200_test_layernorm_bwd(const void* g, const float* x, const float* stats, const float* gamma, const float* dx_res, float* dx_out,
                            void* dx_cast, float* dgamma, float* dbeta, int M, int H, int bf16_mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  return bf16_mode ? test_layernorm_bwd_t<bf16>(g, x, stats, gamma, dx_res, dx_out, dx_cast, dgamma, dbeta, M, H, st)
                   : test_layernorm_bwd_t<float>(g, x, stats, gamma, dx_res, dx_out, dx_cast, dgamma, dbeta, M, H, st);
}
}  // extern "C"
template <class T>
static int test_instnorm_bwd_t(int two, const void* dout, const void* act, const void* ra, const float* mra, const void* rb, const float* mrb,
                               int N, int C, long V, double* acc, void* da, void* db, cudaStream_t st) {
  typedef typename RawOf<T>::type TR;
  constexpr int VN = Vec16<T>::N;
  B200_CHECK(C % VN == 0 && 256 % (C / VN) == 0, "InstanceNorm channel count %d unsupported", C);
  ClView pv{C, 0};
  const size_t red_smem = 256 * 3 * VN * sizeof(float), cst_smem = 7 * (size_t)C * sizeof(float);
  dim3 gr(in_grid_x(V, C / VN), N), ga(in_grid_x(V, C / VN) * 2, N);
  B200_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 3 * N * C, st));
  if (two) {
    if (act) {
      B200_CUDA(launch_pdl(in_bwd_reduce_kernel<T, true, false>, gr, dim3(256), red_smem, st, (const T*)dout, pv, (const T*)act, pv, (const TR*)ra, pv, (const TR*)rb, pv, C, V, acc, mra, mrb));
      B200_LAUNCH_CHECK();
      B200_CUDA(launch_pdl(in_bwd_apply_kernel<T, true, false>, ga, dim3(256), cst_smem, st, (const T*)dout, pv, (const T*)act, pv, (const TR*)ra, pv, mra, (const TR*)rb, pv, mrb,
                           C, V, (const double*)acc, (T*)da, pv, (T*)db, pv));
      B200_LAUNCH_CHECK();
    } else {   // the engine's default: the sign is recomputed from the raw conv outputs
      B200_CUDA(launch_pdl(in_bwd_reduce_kernel<T, true, true>, gr, dim3(256), red_smem, st, (const T*)dout, pv, (const T*)act, pv, (const TR*)ra, pv, (const TR*)rb, pv, C, V, acc, mra, mrb));
      B200_LAUNCH_CHECK();
      B200_CUDA(launch_pdl(in_bwd_apply_kernel<T, true, true>, ga, dim3(256), cst_smem, st, (const T*)dout, pv, (const T*)act, pv, (const TR*)ra, pv, mra, (const TR*)rb, pv, mrb,
                           C, V, (const double*)acc, (T*)da, pv, (T*)db, pv));
      B200_LAUNCH_CHECK();
    }
  } else {
    B200_CUDA(launch_pdl(in_bwd_reduce_kernel<T, false>, gr, dim3(256), red_smem, st, (const T*)dout, pv, (const T*)act, pv, (const TR*)nullptr, pv, (const TR*)nullptr, pv, C, V, acc, (const float*)nullptr, (const float*)nullptr));
    B200_LAUNCH_CHECK();
    B200_CUDA(launch_pdl(in_bwd_apply_kernel<T, false>, ga, dim3(256), cst_smem, st, (const T*)dout, pv, (const T*)act, pv, (const TR*)nullptr, pv, mra, (const TR*)nullptr, pv,
                         (const float*)nullptr, C, V, (const double*)acc, (T*)da, pv, (T*)nullptr, pv));
    B200_LAUNCH_CHECK();
  }
  return 0;
}
extern "C" {
/* InstanceNorm(+LeakyReLU) backward of the residual conv block (exec.cuh res_bwd), channels-last [N,V,C]:
 *  two != 0: out = lrelu(norm(c2) + norm(c3)): dout, act = out (T; NULL: the sign is recomputed from c2, c3), ra = c2, rb = c3 (raw conv outputs: fp16 in bf16 mode), mra / mrb
 *            = (mean, rstd) [N][C][2] -> da = d c2, db = d c3;
 *  two == 0: act = lrelu(norm(c1)) (T; the normalised value is recovered from it), mra = (mean, rstd) of c1 -> da = d c1.
 * acc: double [N][C][3] scratch */
int b200_test_instnorm_bwd(int two, const void* dout, const void* act, const void* ra, const float* mra, const void* rb, const float* mrb,
                           int N, int C, int64_t V, double* acc, void* da, void* db, int bf16_mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  return bf16_mode ? test_instnorm_bwd_t<bf16>(two, dout, act, ra, mra, rb, mrb, N, C, V, acc, da, db, st)
                   : test_instnorm_bwd_t<float>(two, dout, act, ra, mra, rb, mrb, N, C, V, acc, da, db, st);
}
}  // extern "C"
template <class T, int CO>
static int test_head_bwd_t(const float* dlogits, const void* d0, const float* Wh, int ncls, int N, long V, void* g, float* dWh, float* dbh, cudaStream_t st) {
  const long chunk = 2048;
  dim3 grid((unsigned)((V + chunk - 1) / chunk), N);
  B200_CUDA(cudaFuncSetAttribute(head_bwd2_kernel<T, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  B200_CUDA(cudaMemsetAsync(dWh, 0, sizeof(float) * ncls * CO, st));
  B200_CUDA(cudaMemsetAsync(dbh, 0, sizeof(float) * ncls, st));
  head_bwd2_kernel<T, CO><<<grid, 256, head_bwd2_smem(CO), st>>>(dlogits, (const T*)d0, Wh, ncls, V, chunk, (T*)g, dWh, dbh);
  B200_LAUNCH_CHECK();
  return 0;
}
extern "C" {
/* 1x1x1 head backward (UnetOutBlock, unetr.py:175): dlogits fp32 [N][ncls][V], d0 [N,V,fs] (T) -> g = d(d0) [N,V,fs] (T), dWh [ncls][fs],
 * dbh [ncls]; fs in {8,16,32}, ncls <= 16 */
int b200_test_head_bwd(const float* dlogits, const void* d0, const float* Wh, int ncls, int fs, int N, int64_t V, void* g, float* dWh,
                       float* dbh, int bf16_mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B200_CHECK((fs == 8 || fs == 16 || fs == 32) && ncls >= 1 && ncls <= 16, "b200_test_head_bwd: fs in {8,16,32}, ncls <= 16");
  if (bf16_mode) return fs == 8 ? test_head_bwd_t<bf16, 8>(dlogits, d0, Wh, ncls, N, V, g, dWh, dbh, st) : fs == 16 ? test_head_bwd_t<bf16, 16>(dlogits, d0, Wh, ncls, N, V, g, dWh, dbh, st) : test_head_bwd_t<bf16, 32>(dlogits, d0, Wh, ncls, N, V, g, dWh, dbh, st);
  return fs == 8 ? test_head_bwd_t<float, 8>(dlogits, d0, Wh, ncls, N, V, g, dWh, dbh, st) : fs == 16 ? test_head_bwd_t<float, 16>(dlogits, d0, Wh, ncls, N, V, g, dWh, dbh, st) : test_head_bwd_t<float, 32>(dlogits, d0, Wh, ncls, N, V, g, dWh, dbh, st);
}

void b200_test_set_debug_buffer(void* dev_ptr) { tc::g_dbg = (long long*)dev_ptr; }

int b200_test_tc_gemm(const void* a, const void* b, float* out, int M, int N, int K, int a_mn, int b_mn, void* stream) {
  return tc_gemm_test((const bf16*)a, (const bf16*)b, out, M, N, K, a_mn, b_mn, (cudaStream_t)stream);
}

int b200_test_tc_gemm_grouped(const void* const* a, const void* const* b, float* const* out, const int* M, const int* N, const int* K, int n, int mn,
                              void* stream) {
  return tc_gemm_grouped_test((const bf16* const*)a, (const bf16* const*)b, out, M, N, K, n, mn, (cudaStream_t)stream);
}

// test hook for the fused attention forward: qkv [B*L][3H] bf16 -> probs [B][nh][L][Lp] bf16 (may be NULL), att [B*L][H] bf16
int b200_test_tc_attention(const void* qkv, void* probs, void* att, int B, int heads, int L, int Lp, int H, float scale, void* stream) {
  B200_CHECK(tc::attention_fused_supported(L, Lp, H, heads), "fused attention needs head_dim 64, 16 <= L <= 256, Lp %% 8 == 0");
  return tc::attention_fused_fwd((const bf16*)qkv, (bf16*)probs, (bf16*)att, B, heads, L, Lp, H, scale, (cudaStream_t)stream);
}

// query-row half of the attention backward: probs/datt in, dS [B][heads][L][Lp] and the Q third of dqkv [B*L][3H] out
int b200_test_tc_attention_bwd(const void* qkv, const void* probs, const void* datt, void* dS, void* dqkv, int B, int heads, int L, int Lp,
                               int H, float scale, void* stream) {
  B200_CHECK(tc::attention_fused_supported(L, Lp, H, heads), "fused attention needs head_dim 64, 16 <= L <= 256, Lp %% 8 == 0");
  return tc::attention_fused_bwd_dq((const bf16*)qkv, (const bf16*)probs, (const bf16*)datt, (bf16*)dS, (bf16*)dqkv, B, heads, L, Lp, H, scale,
                                    (cudaStream_t)stream);
}

// key-row half: dV = probs^T dO -> V third of dqkv, dK = dS^T Q -> K third (the Q third is not touched)
int b200_test_tc_attention_bwd_kv(const void* qkv, const void* probs, const void* dS, const void* datt, void* dqkv, int B, int heads, int L,
                                  int Lp, int H, void* stream) {
  B200_CHECK(tc::attention_fused_supported(L, Lp, H, heads), "fused attention needs head_dim 64, 16 <= L <= 256, Lp %% 8 == 0");
  return tc::attention_fused_bwd_kv((const bf16*)qkv, (const bf16*)probs, (const bf16*)dS, (const bf16*)datt, (bf16*)dqkv, B, heads, L, Lp, H,
                                    (cudaStream_t)stream);
}

}  // extern "C"
