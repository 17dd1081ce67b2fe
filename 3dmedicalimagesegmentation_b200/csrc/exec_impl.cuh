// ExecIface adaptor over Exec<T>; included by exec_f32.cu / exec_bf16.cu only.
#pragma once
#include "exec.cuh"

namespace b200 {
template <class T>
struct ExecImpl : ExecIface {
  Exec<T> e;
  explicit ExecImpl(const UnetrConfig& c) : e(c) {}
  size_t workspace_bytes(bool with_backward) override { e.layout(nullptr, with_backward); return e.w.bytes; }
  int forward(const float* const* P, const float* x, char* ws, float* enc4_out, float* logits_out, int flags, cudaStream_t st) override {
    return e.forward(P, x, ws, enc4_out, logits_out, flags, st);
  }
  int backward(const float* const* P, float* const* G, const float* x, char* ws, const float* d_enc4, const float* d_logits, int flags,
               cudaStream_t st) override {
    return e.backward(P, G, x, ws, d_enc4, d_logits, flags, st);
  }
  const void* peek(const char* name, size_t* bytes) override { return e.peek(name, bytes); }
  size_t packed_bytes() override { return e.layout_packed(nullptr); }
  void set_packed(char* buf) override { e.packed_base = buf; }
  long long packed_cast_offset(int pidx) override { return e.packed_cast_offset(pidx); }
  int pack_convs(const float* const* P, char* packed, cudaStream_t st) override {
    if (!Exec<T>::kTC) return 0;
    if (!packed) { set_error("b200_unetr_pack_convs: no packed-weight buffer"); return 1; }
    e.layout_packed(packed);
    return e.pack_weights(P, st, false);
  }
  void set_grad_events(cudaEvent_t* ev, int n) override { e.n_grad_ev = (n == 4 || n == 7 || n == 13) ? n : 0; for (int i = 0; i < e.n_grad_ev; ++i) e.grad_ev[i] = ev[i]; }
};
}  // namespace b200
