/* C ABI of libunetr_b200.so -- the B200 (sm_100a) implementation of the UNETR hot path of
 * ilkyyldz95/3DmedicalImageSegmentation.
 *
 * The reference has no FFI of its own: its boundary is a Python nn.Module plus three callables
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it stands behind.
 * Conventions: plain device pointers and sizes only (no torch types); the caller owns every buffer;
 * every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns 0 on success
 * or non-zero with a message retrievable by b200_last_error().  Nothing here ever computes on the CPU.
 *
 * Parameter / gradient tables: arrays of B200_PARAM_COUNT fp32 device pointers in PyTorch-native
 * layouts, ordered as enum ParamIdx in csrc/exec.cuh (== the reference state-dict order of
 * SURVEY 8b without the unused cls_token).  A null gradient pointer means "not wanted"
 * (the parameter's .grad stays None, as in the reference's ranking stages).
 */
#ifndef UNETR_B200_H
#define UNETR_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B200_PARAM_COUNT 164

/* flags for forward/backward */
#define B200_NEED_ENCODER_GRAD 1 /* 0 == forward(x, freeze_encoder=True), unetr.py:183-192 */
#define B200_HAS_DLOGITS 2
#define B200_HAS_DENC4 4
#define B200_INPLACE_WGRADS 32 /* backward with gradient events set: launch the conv-stack weight gradients in place (otherwise: after the ViT backward, event 0 last) */
#define B200_NO_BACKWARD 16 /* forward only: no backward call follows (inference); buffers only the backward reads may be left unwritten */
#define B200_WEIGHTS_PACKED 64 /* forward only: the bf16 weight copies in the packed buffer are current (b200_unetr_set_packed_weights) */

typedef struct {
  int32_t batch, in_channels, out_channels; /* unetr.py:29-30 */
  int32_t img0, img1, img2;                 /* img_size, unetr.py:31; multiples of 16 */
  int32_t feature_size, hidden_size, mlp_dim, num_heads; /* unetr.py:32-35 */
  int32_t conv_patch_embed;                 /* pos_embed == "conv" (unetr.py:36,66) */
  int32_t mode;                             /* 0 = fp32 parity mode, 1 = bf16 throughput mode */
} b200_unetr_config;

const char* b200_last_error(void);
/* 0 when the current device is an sm_100 part; the Python layer refuses to run otherwise (no fallback) */
int b200_device_check(void);

/* ---- UNETR.forward / autograd backward  (unetr.py:182-208; monai.networks.nets.UNETR at seg:36,221) ---- */
void* b200_unetr_create(const b200_unetr_config* cfg);
void b200_unetr_destroy(void* handle);
size_t b200_unetr_workspace_bytes(void* handle, int with_backward);
/* x: [B,Cin,S0,S1,S2] fp32 NCDHW.  enc4_out: [B,8*fs,S/8...] fp32 or NULL.  logits_out: [B,ncls,S...] fp32 or NULL. */
int b200_unetr_forward(void* handle, const float* const* params, const float* x, void* workspace, float* enc4_out,
                       float* logits_out, int flags, void* stream);
/* consumes the workspace filled by the matching forward; writes (not accumulates) every non-null gradient */
int b200_unetr_backward(void* handle, const float* const* params, float* const* grads, const float* x, void* workspace,
                        const float* d_enc4, const float* d_logits, int flags, void* stream);

/* ---- monai.losses.DiceCELoss(to_onehot_y=True, softmax=True)  (seg:404,222) ---- */
size_t b200_dicece_scratch_bytes(int batch, int classes);
/* out3 = {loss, dice term, ce term}; labels: [B,1,V] fp32 holding class ids */
int b200_dicece_forward(const float* logits, const float* labels, int batch, int classes, int64_t voxels, void* scratch,
                        float* out3, void* stream);
int b200_dicece_backward(const float* logits, const float* labels, int batch, int classes, int64_t voxels,
                         const void* scratch, const float* upstream, float* dlogits, void* stream);

/* ---- monai.losses.DiceCELoss(to_onehot_y=False, sigmoid=True)  (seg:480) ----
 * target: [B,C,V] fp32 multi-hot (seg:65-93).  Dice on sigmoid(logits); CE against argmax_c(target) (MONAI 0.6.0 rule for a
 * target with as many channels as the prediction).  Scratch size/layout as b200_dicece_scratch_bytes; C <= 16. */
int b200_dicece_sigmoid_forward(const float* logits, const float* target, int batch, int channels, int64_t voxels,
                                void* scratch, float* out3, void* stream);
int b200_dicece_sigmoid_backward(const float* logits, const float* target, int batch, int channels, int64_t voxels,
                                 const void* scratch, const float* upstream, float* dlogits, void* stream);

/* ---- validation tail: monai.metrics.DiceMetric / ConfusionMatrixMetric (seg:485-494, used seg:110-126,153-188; 8f N2) ----
 * counts: [B][C][3] doubles = (|y & p|, |p|, |y|), zeroed by the call.
 *   _onehot: y_pred, y = [B,C,V] fp32 one-hot tensors as the reference passes them (post_pred / post_label, seg:405-406)
 *   _labels: mask = [B,V] uint8 argmax (b200_sw_finalize), labels = [B,V] fp32 holding class ids; C <= 32 */
int b200_seg_counts_onehot(const float* y_pred, const float* y, int batch, int classes, int64_t voxels, double* counts,
                           void* stream);
int b200_seg_counts_labels(const uint8_t* mask, const float* labels, int batch, int classes, int64_t voxels,
                           double* counts, void* stream);
/* dice[n][c] = 2|y&p|/(|y|+|p|), NaN when |y| == 0; confusion[n][c] = (tp, fp, tn, fn).  Either output may be NULL. */
int b200_seg_metrics(const double* counts, int rows, int classes, int64_t voxels, float* dice, float* confusion,
                     void* stream);
/* MONAI do_metric_reduction over f[N][C][K]: reduction 0 = "mean" -> out[K], 1 = "mean_batch" -> out[C][K]; NaN-aware;
 * not_nans (same shape as out) may be NULL */
int b200_metric_reduce(const float* f, int n, int classes, int k, int reduction, float* out, float* not_nans, void* stream);
/* rows of (tp, fp, tn, fn) -> metric 0 precision, 1 sensitivity; NaN where the denominator is 0 */
int b200_confusion_metric(const float* confusion, int rows, int metric, float* out, void* stream);

/* ---- extract_triplets_more_partitions + BTLoss  (rank:59-133, rank:202-217) ----
 * 4 samples (batch1[0], batch1[1], batch2[0], batch2[1]), one slice index per partition along the sliced axis. */
typedef struct {
  const float* src[4];
  float* grad[4];             /* same geometry as src; must be zero-filled by the caller */
  int64_t stride_c, stride_slice, stride_f0, stride_f1; /* element strides */
  int32_t channels, f0, f1;
  int32_t idx[4];
  float temperature;          /* rank:327 */
} b200_rank_geom;
size_t b200_ranking_scratch_bytes(int channels);
int b200_ranking_forward(const b200_rank_geom* g, void* scratch, float* loss_out, void* stream);
int b200_ranking_backward(const b200_rank_geom* g, const void* scratch, const float* upstream, void* stream);

/* ---- monai.inferers.sliding_window_inference, constant blending  (seg:109,143,694) ---- */
typedef struct {
  int32_t channels, d, h, w, pad_d, pad_h, pad_w, padded_d, padded_h, padded_w, roi0, roi1, roi2;
} b200_sw_geom;
/* starts: n x {batch index, s0, s1, s2} (n <= 16) */
int b200_sw_gather(const float* volume, float* windows, const b200_sw_geom* g, const int32_t* starts, int n, float cval,
                   void* stream);
int b200_sw_accumulate(float* acc, const float* pred, const b200_sw_geom* g, const int32_t* start4, void* stream);
/* the same overlap-add for n <= 16 windows of one batch item in one launch (pred: [n,C,roi]); per-voxel addition order =
 * window order, so the result is bit-identical to n calls of b200_sw_accumulate */
int b200_sw_accumulate_n(float* acc, const float* pred, const b200_sw_geom* g, const int32_t* starts, int n, void* stream);
/* per-axis window starts (n0,n1,n2 <= 64).  out and/or mask (uint8 argmax) may be NULL. */
int b200_sw_finalize(const float* acc, float* out, uint8_t* mask, const b200_sw_geom* g, int batch, const int32_t* s0,
                     int n0, const int32_t* s1, int n1, const int32_t* s2, int n2, void* stream);

/* same pass with the validation tail fused: labels [B,D,H,W] fp32 class ids, counts [B][C][3] as b200_seg_counts_labels
 * (zeroed by the call); both NULL = plain b200_sw_finalize */
int b200_sw_finalize_metric(const float* acc, float* out, uint8_t* mask, const b200_sw_geom* g, int batch,
                            const int32_t* s0, int n0, const int32_t* s1, int n1, const int32_t* s2, int n2,
                            const float* labels, double* counts, void* stream);

/* ---- slab-owned sliding window (multi-GPU, one process per GPU).  The window list of a volume is cut into contiguous
 * chunks per rank; rank r owns an equal slab of the padded rows.
 * Contributions that fall into another rank's rows travel there as row-clipped "pieces" (b200_sw_pack_rows -> ncclSend/Recv) and the
 * owner adds all pieces of a voxel in global window order (b200_sw_accumulate_slab), so the result is bit-identical to the
 * single-GPU loop of b200_sw_accumulate_n.  pieces: n x {s0, s1, s2, x_lo, x_hi, nx, xbase} int32 -- window start, the padded rows
 * [x_lo, x_hi) this piece covers, and its buffer preds[i] = [C][nx][roi1][roi2] whose row 0 is padded row xbase.  acc is the slab
 * [C][nrows][padded_h][padded_w] of padded rows [xoff, xoff + nrows).  preds is a HOST array of n device pointers. */
int b200_sw_accumulate_slab(float* acc, const b200_sw_geom* g, const void* const* preds, const int32_t* pieces, int n, int xoff, int nrows,
                            void* stream);
int b200_sw_pack_rows(const float* pred, float* dst, const b200_sw_geom* g, int x_from, int nx, void* stream);
/* normalise pass of a slab: un-padded rows [d0, d0 + nd) of batch item `item`; out (nullable) holds un-padded rows
 * [out_d0, out_d0 + out_rows) of that item; mask / labels / counts are the FULL-size tensors of b200_sw_finalize_metric (counts are
 * accumulated, not zeroed). */
int b200_sw_finalize_slab(const float* acc, float* out, uint8_t* mask, const b200_sw_geom* g, int item, const int32_t* s0, int n0,
                          const int32_t* s1, int n1, const int32_t* s2, int n2, const float* labels, double* counts,
                          int d0, int nd, int acc_xoff, int acc_rows, int out_d0, int out_rows, void* stream);

/* ---- GPU-side crop sampling and augmentation on a volume resident in device memory: the per-iteration tail of the
 *      reference's training transforms -- RandCropByPosNegLabeld (seg:341-350), RandFlipd x3 / RandRotate90d / RandShiftIntensityd
 *      (seg:351-375), RandSpatialCropSamplesd (rank:365-369), ConvertToMultiChannelBasedOnBratsClassesd (seg:65-93).  The random
 *      draws are made on the host (numpy RandomState streams, as MONAI's Randomizable does); the kernels do everything that touches
 *      voxels.  image [Ci][D][H][W], label [Cl][D][H][W] fp32.
 *   b200_aug_index        once per volume: per-block counts + exclusive prefix of the foreground (any label channel != 0) and
 *                         background (not foreground, and any image channel > threshold; image NULL: not foreground) voxel sets of
 *                         monai.transforms.utils.map_binary_to_indices -- never materialised as index lists; totals = their sizes
 *   b200_aug_pick_centers picks[i] = {use_fg, k}: crop start of "the k-th voxel of that set in raster order" (= fg_indices[k]) after
 *                         correct_crop_centers, written to device memory (no host round trip before the crop)
 *   b200_aug_crop         n crops in one gather launch; maps[i]: output axis a reads crop axis perm[a], reversed when flip[a] (the
 *                         composition of the flips and rot90 the host drew), image + shift; brats != 0: label map -> 4 multi-hot
 *                         channels (background, TC, WT, ET) */
typedef struct { int32_t perm[3]; int32_t flip[3]; float shift; } b200_aug_map;
int b200_aug_blocks(int64_t voxels);
int b200_aug_index(const float* label, int label_channels, const float* image, int image_channels, float image_threshold, int64_t voxels,
                   int32_t* prefix, int64_t* totals, void* stream);
int b200_aug_pick_centers(const float* label, int label_channels, const float* image, int image_channels, float image_threshold, int d, int h,
                          int w, const int32_t* prefix, const int64_t* picks, int n, int roi0, int roi1, int roi2, int32_t* starts, void* stream);
int b200_aug_crop(const float* image, int image_channels, const float* label, int label_channels, int d, int h, int w, const int32_t* starts,
                  const b200_aug_map* maps, int n, int roi0, int roi1, int roi2, int brats, float* out_image, float* out_label, void* stream);

/* debug: copy a named workspace buffer of the last forward/backward (device to device, synchronous) */
int b200_unetr_peek(void* handle, const char* name, void* dst, size_t cap);

/* ---- instrumentation: kernels launched so far; per-op CUDA-event timing (tags = layer types) ---- */
unsigned long long b200_launch_count(void);
void b200_prof_enable(int on);
int b200_prof_report(char* buf, int cap);

/* ---- packed bf16 weight copies (bf16 mode).  The tcgen05 engine reads bf16 copies of the GEMM / convolution weights in its own
 *      layouts from ONE caller-owned device buffer of b200_unetr_packed_bytes() bytes (independent of the batch size; handles of the
 *      same network may share it).  b200_unetr_forward refreshes every copy from the fp32 parameters unless B200_WEIGHTS_PACKED says
 *      they are current.  An optimizer can keep them current instead: b200_adamw_step writes the PLAIN casts (byte offset from
 *      b200_unetr_packed_cast_offset, -1 = the parameter has none) while it updates the parameters, and b200_unetr_pack_convs
 *      refreshes the re-laid-out convolution / transposed-convolution copies (one launch, 7 M parameters at configs[1]). */
size_t b200_unetr_packed_bytes(void* handle);
void b200_unetr_set_packed_weights(void* handle, void* buffer);
int64_t b200_unetr_packed_cast_offset(void* handle, int param_index);
int b200_unetr_pack_convs(void* handle, const float* const* params, void* packed, void* stream);

/* ---- fused multi-tensor AdamW (stands behind torch.optim.AdamW: unetr_segmentation_3d.py:522,225-226;
 *      unetr_ranking_pretraining_3d.py:466,214-215).  tensors: device array of {float* p; const float* g; float* m; float* v;
 *      int64 n; bf16* s0} -- s0 (nullable) is the plain bf16 mirror of the parameter in the packed-weight buffer
 *      (b200_unetr_packed_cast_offset), which the same launch keeps current; chunks: device array of {int32 tensor; int32 pad; int64 start}, each covering b200_adamw_chunk() elements.
 *      Parameters without a gradient are left out of the table (their state does not move).  step is 1-based. */
long b200_adamw_chunk(void);
int b200_adamw_step(const void* tensors, const void* chunks, int n_chunks, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, void* stream);
/* dev_step != NULL: the 1-based update count is the device int *dev_step, advanced by one on the stream before the update when
 * step > 0 (step == 0: a further launch of the same update over other chunks; the count is read, not advanced) -- every step is
 * then the same launch sequence and can be replayed from a CUDA graph. */
int b200_adamw_step_capturable(const void* tensors, const void* chunks, int n_chunks, float lr, float beta1, float beta2, float eps,
                               float weight_decay, int step, int* dev_step, void* stream);

/* Optional gradient-ready events for data-parallel overlap: cudaEvent_t handles recorded inside b200_unetr_backward when a group
 * of parameter gradients is final -- [0] conv encoders/decoders + head, then groups of transformer blocks from the top; the first
 * of those also covers vit.norm, the last also the patch embedding.  n = 4: groups of 4 blocks ([1] 8..11, [2] 4..7, [3] 0..3);
 * n = 7: groups of 2 blocks; n = 13: one block per event.  The deferred weight-gradient launches use the same grouping.  n = 0
 * switches the recording off.  Only a full backward (logits gradient + trainable encoder) records all of them. */
void b200_unetr_set_grad_events(void* handle, void* const* events, int n);

/* ---- op-level test hooks (parity tests of single kernels; not part of the reference surface) ---- */
/* D[M,N] = A[M,K] * B[N,K]^T with bf16 operands on the tcgen05 engine; a_mn/b_mn select MN-major operands
 * (A stored [K,M] / B stored [K,N]).  out fp32. */
/* tcgen05 implicit-GEMM conv3d (k=1|3, same padding) on channels-last bf16; see csrc/capi.cu for the argument layout */
int b200_test_tc_conv(const void* x, int in_pitch, int in_coff, int Ci, int N, int D, int H, int W, const float* w, int Co, int ks,
                      void* out, int out_pitch, int out_coff, int accumulate, int dgrad, double* stats, void* scratch, void* stream);
/* fused 3^3 + 1^3 conv of the residual block: mode 1 = two outputs of one input (forward), 2 = two inputs of one output (dgrad) */
int b200_test_tc_conv_fused(const void* x, const void* x2, int Ci, int Co, int N, int D, int H, int W, const float* w3, const float* w1, int mode,
                            void* out, void* out2, double* stats, double* stats2, void* scratch, void* stream);
int b200_test_tc_wgrad(const void* x, int x_pitch, int x_coff, int Ci, const void* dy, int dy_pitch, int dy_coff, int Co, int N, int D, int H,
                       int W, int ks, float* dW, void* stream);
/* op-level hooks for the element-wise kernels of the backward, each pinned on its own against fp32 torch with an injected upstream
 * gradient (tests/test_gpu_ops_bwd.py).  bf16_mode != 0: activations / gradients are bf16 (raw conv outputs fp16), else fp32.
 *  layernorm_bwd: g [M,H], x fp32 [M,H], stats fp32 [M][2] = (mean, rstd), gamma [H]; dx_out fp32 = dx_res (nullable) + dL/dx,
 *                 dx_cast = the same in the activation type, dgamma / dbeta [H]                       (transformer blocks, unetr.py:69-76)
 *  instnorm_bwd:  channels-last [N,V,C]; two != 0: out = lrelu(norm(c2) + norm(c3)) with ra = c2, rb = c3 -> da = d c2, db = d c3 (act NULL: sign recomputed from c2, c3);
 *                 two == 0: act = lrelu(norm(c1)) -> da = d c1; mra / mrb = (mean, rstd) [N][C][2]; acc: double [N][C][3] scratch
 *  head_bwd:      dlogits fp32 [N][ncls][V], d0 [N,V,fs] -> g = d(d0), dWh [ncls][fs], dbh [ncls]      (UnetOutBlock, unetr.py:175) */
int b200_test_layernorm_bwd(const void* g, const float* x, const float* stats, const float* gamma, const float* dx_res, float* dx_out,
                            void* dx_cast, float* dgamma, float* dbeta, int M, int H, int bf16_mode, void* stream);
int b200_test_instnorm_bwd(int two, const void* dout, const void* act, const void* ra, const float* mra, const void* rb, const float* mrb,
                           int N, int C, int64_t V, double* acc, void* da, void* db, int bf16_mode, void* stream);
int b200_test_head_bwd(const float* dlogits, const void* d0, const float* Wh, int ncls, int fs, int N, int64_t V, void* g, float* dWh,
                       float* dbh, int bf16_mode, void* stream);
/* dgrad of a 3^3 convolution with the first pass of the InstanceNorm + LeakyReLU backward folded into its epilogue (res_bwd, exec.cuh):
 * dy [N,D,H,W,Co] bf16, w fp32 [Co][Ci][27], act = the saved activation lrelu(norm(.)) [N,D,H,W,Ci] bf16 -> out = dx bf16, acc double
 * [N][Ci][3] = (sum g, sum g*n, -); *folded = 0 when the kernel that ran does not fold (acc untouched).  scratch: 2*Co*Ci*27 bf16 */
int b200_test_tc_conv_dgrad_normbwd(const void* dy, int Co, int Ci, int N, int D, int H, int W, const float* w, const void* act, void* out,
                                    double* acc, int* folded, void* scratch, void* stream);
/* tuning aid: 16 x int64 device buffer receiving CTA-0 clock64 phase stamps of the next tcgen05 GEMM launches (NULL = off) */
void b200_test_set_debug_buffer(void* dev_ptr);
/* in-situ trace of the tcgen05 launches of the following calls: buf = device int64[2*cap] pre-filled with (INT64_MAX, 0)
 * pairs receives (min CTA start, max CTA end) of %globaltimer per launch; b200_trace_tags lists one "kind dims" line per slot */
void b200_trace_begin(void* buf, int cap);
int b200_trace_count(void);
int b200_trace_tags(char* out, int cap);
int b200_test_tc_gemm(const void* a, const void* b, float* out, int M, int N, int K, int a_mn, int b_mn, void* stream);
/* op-level hook for the grouped tcgen05 GEMM that forms the deferred ViT weight gradients (dW = dY^T X of
 * /root/reference/unetr.py:69-76's 12 x {qkv, out_proj, linear1, linear2}) in one persistent launch: n independent problems
 * out_i [M_i, N_i] fp32 = A_i B_i^T; mn != 0: a_i is [K_i, M_i] and b_i is [K_i, N_i] bf16 (MN-major, the weight-gradient form),
 * mn == 0: a_i [M_i, K_i], b_i [N_i, K_i].  a / b / out are HOST arrays of n device pointers, M / N / K host arrays of n ints. */
int b200_test_tc_gemm_grouped(const void* const* a, const void* const* b, float* const* out, const int* M, const int* N, const int* K,
                              int n, int mn, void* stream);
/* op-level hook for the fused attention forward (SABlock): qkv [B*L][3H] bf16 with columns [Q|K|V] x head x 64;
 * probs [B][heads][L][Lp] bf16 (softmax(scale QK^T), may be NULL); att [B*L][H] bf16.  head_dim 64, 16 <= L <= 256. */
int b200_test_tc_attention(const void* qkv, void* probs, void* att, int batch, int heads, int L, int Lp, int H, float scale,
                           void* stream);
/* query-row half of its backward: dS = probs * (dO V^T - rowsum(dO V^T * probs)) * scale -> dS [B][heads][L][Lp] bf16, and
 * dQ = dS K -> columns [h*64, h*64+64) of dqkv [B*L][3H] bf16 (the K and V thirds are not touched) */
int b200_test_tc_attention_bwd(const void* qkv, const void* probs, const void* datt, void* dS, void* dqkv, int batch, int heads,
                               int L, int Lp, int H, float scale, void* stream);
/* key-row half: dV = probs^T dO -> columns [2H + h*64, +64) of dqkv, dK = dS^T Q -> columns [H + h*64, +64) */
int b200_test_tc_attention_bwd_kv(const void* qkv, const void* probs, const void* dS, const void* datt, void* dqkv, int batch,
                                  int heads, int L, int Lp, int H, void* stream);

#ifdef __cplusplus
}
#endif
#endif
