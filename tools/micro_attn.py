"""Phase stamps (clock64, CTA 0) and event timing of the fused attention forward at the configs[1] / sliding-window shapes."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("3dmedicalimagesegmentation_b200")
L_ = pkg._lib; lib = L_.load(); dev = torch.device("cuda", 0)
for B in (2, 4):
    heads, L, H = 12, 216, 768
    qkv = torch.randn(B * L, 3 * H, device=dev).to(torch.bfloat16)
    probs = torch.empty(B, heads, L, L, dtype=torch.bfloat16, device=dev)
    att = torch.empty(B * L, H, dtype=torch.bfloat16, device=dev)
    dbg = torch.zeros(1024, dtype=torch.int64, device=dev)
    run = lambda pr: L_.check(lib.b200_test_tc_attention(L_.ptr(qkv), L_.ptr(pr) if pr is not None else None, L_.ptr(att), B, heads, L, L, H, 0.125, L_.stream_ptr()), "att")
    for _ in range(3): run(probs)
    lib.b200_test_set_debug_buffer(L_.ptr(dbg)); run(probs); torch.cuda.synchronize(); lib.b200_test_set_debug_buffer(None)
    d = dbg[:8].tolist()
    names = ["start", "qk_landed", "S_ready", "pass1_done", "P_published", "O_ready", "end", "v_landed"]
    print(f"B={B}: " + "  ".join(f"{n} {d[i] - d[0]}" for i, n in enumerate(names)))
    for pr, tag in ((probs, "with P store"), (None, "no P store")):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): run(pr)
        e1.record(); torch.cuda.synchronize()
        print(f"   {tag}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per launch (back to back, PDL)")

# backward kernels at the configs[1] shape (B = 2): query-row half (attn_kernel<1>) and key-row half (attn_bwd_kv_kernel)
B, heads, L, H = 2, 12, 216, 768
qkv = torch.randn(B * L, 3 * H, device=dev).to(torch.bfloat16)
datt = torch.randn(B * L, H, device=dev).to(torch.bfloat16)
probs = torch.softmax(torch.randn(B, heads, L, L, device=dev), -1).to(torch.bfloat16)
dS = torch.empty_like(probs); dqkv = torch.empty(B * L, 3 * H, dtype=torch.bfloat16, device=dev)
def bwd():
    L_.check(lib.b200_test_tc_attention_bwd(L_.ptr(qkv), L_.ptr(probs), L_.ptr(datt), L_.ptr(dS), L_.ptr(dqkv), B, heads, L, L, H, 0.125, L_.stream_ptr()), "bwd")
    L_.check(lib.b200_test_tc_attention_bwd_kv(L_.ptr(qkv), L_.ptr(probs), L_.ptr(dS), L_.ptr(datt), L_.ptr(dqkv), B, heads, L, L, H, L_.stream_ptr()), "bwd_kv")
for _ in range(3): bwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): bwd()
e1.record(); torch.cuda.synchronize()
print(f"backward (dS/dQ kernel + dV/dK kernel): {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per pair (back to back, PDL)")
