"""Development probe: where the roles of one CTA of the fused backward wait (cycles per pair).  Needs a library
built with -DFA_BWD_TRACE selected through FA_B200_LIB (ab_trace.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa

L = fa.lib()
L.fa_debug_set_prof_buffer.argtypes = [ctypes.c_void_p]
B, H, n, d, causal = 1, 16, 16384, int(sys.argv[1]) if len(sys.argv) > 1 else 128, int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator(device="cuda").manual_seed(5)
Q, K, V, dO = (torch.rand((B, H, n, d), device="cuda", generator=g).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(4))
O = torch.empty_like(Q); Ls = torch.empty((B, H, n), device="cuda")
scale = d ** -0.5
fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, Ls, causal, B, H, fa.BF16)
dQ, dK, dV = (torch.empty((B, H, n, d), device="cuda") for _ in range(3))
wsb = fa.workspace_bytes_backward(n, d, B, H); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
run = lambda: fa.flash_attention_backward(Q, K, V, O, dO, Ls, dQ, dK, dV, n, d, scale, H * n * d, n * d, causal, B, H, fa.BF16, ws, wsb)
for _ in range(3): run()
prof = torch.zeros(64, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
L.fa_debug_set_prof_buffer(prof.data_ptr()); run(); torch.cuda.synchronize(); L.fa_debug_set_prof_buffer(None)
p = prof.cpu().tolist()
npairs = max(p[63], 1)
per = lambda xs: " ".join(f"{x / npairs:7.0f}" for x in xs)
print(f"d={d} causal={causal}: traced CTA key tile 8 of head 0, {npairs} pairs; cycles PER PAIR")
print("element-wise thread: x_full, phase1, dq_full, stage_free(A), y_full, phase2, stage_free(B), stats-barrier | total")
print("   ", per(p[0:8]), "|", per([p[8]]))
print("MMA issuer: q_full, do_full, p_ready, ds_ready, dq_drained | total")
print("   ", per(p[16:21]), "|", per([p[21]]))
print("dQ reducer: stage_full, ordering counter, read wait, completion wait | total")
print("   ", per(p[32:36]), "|", per([p[36]]))
