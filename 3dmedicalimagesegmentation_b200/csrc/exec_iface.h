// Type-erased handle on Exec<T> so the fp32 and bf16 executors compile in separate translation units (parallel build).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace b200 {

struct UnetrConfig {
  int B, Cin, ncls, S0, S1, S2, fs, hidden, mlp, heads, conv_patch, mode;  // mode 0 = fp32, 1 = bf16
};

struct ExecIface {
  virtual ~ExecIface() {}
  virtual size_t workspace_bytes(bool with_backward) = 0;
  virtual int forward(const float* const* P, const float* x, char* ws, float* enc4_out, float* logits_out, int flags, cudaStream_t st) = 0;
  virtual int backward(const float* const* P, float* const* G, const float* x, char* ws, const float* d_enc4, const float* d_logits, int flags,
                       cudaStream_t st) = 0;
  virtual const void* peek(const char* name, size_t* bytes) = 0;
  // optional: events recorded inside backward() when a group of parameter gradients is final (gradient all-reduce overlap):
  //   [0] convolutional encoders/decoders + head, then groups of transformer blocks from the top: n = 4 -> [1] vit.norm + blocks 8..11,
  //   [2] blocks 4..7, [3] blocks 0..3 + patch embedding; n = 7 -> groups of two blocks; n = 13 -> one block per event
  virtual void set_grad_events(cudaEvent_t* ev, int n) = 0;
  // bf16 mode: caller-owned buffer of packed_bytes() bytes holding the packed bf16 weight copies (see Exec::layout_packed)
  virtual size_t packed_bytes() = 0;
  virtual void set_packed(char* buf) = 0;
  virtual long long packed_cast_offset(int pidx) = 0;
  virtual int pack_convs(const float* const* P, char* packed, cudaStream_t st) = 0;
};
ExecIface* make_exec_f32(const UnetrConfig& c);
ExecIface* make_exec_bf16(const UnetrConfig& c);

}  // namespace b200
