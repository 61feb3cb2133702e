"""Development probe: timeline (SM cycles) of one dQ CTA's hand-offs; needs a library built with
-DFA_BWD_TRACE (make EXTRA=-DFA_BWD_TRACE) selected through FA_B200_LIB."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa
L = fa.lib()
L.fa_debug_set_prof_buffer.argtypes = [ctypes.c_void_p]
B, H, n, d, causal = 1, 16, 16384, 128, False
Q, K, V, dO = (torch.randn((B, H, n, d), device="cuda").to(torch.bfloat16) for _ in range(4))
O = torch.empty_like(Q); Ls = torch.empty((B, H, n), device="cuda")
dQ, dK, dV = (torch.empty((B, H, n, d), device="cuda") for _ in range(3))
wsb = fa.workspace_bytes_backward(n, d, B, H); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
fa.flash_attention_v4_half(Q, K, V, O, n, d, d ** -0.5, H * n * d, n * d, Ls, causal, B, H, fa.BF16)
run = lambda: fa.flash_attention_backward(Q, K, V, O, dO, Ls, dQ, dK, dV, n, d, d ** -0.5, H * n * d, n * d, causal, B, H, fa.BF16, ws, wsb)
run(); run()
prof = torch.zeros(512, dtype=torch.int64, device="cuda")
L.fa_debug_set_prof_buffer(prof.data_ptr()); run(); torch.cuda.synchronize(); L.fa_debug_set_prof_buffer(None)
p = prof.cpu().tolist()[64:64 + 256]
names = {}
for t in (0, 1):
    for k, nm in enumerate(("S ready", "S copied out (X released)", "P done, waiting for dP", "dP ready", "dS part 0 published", "dS part 3 published")):
        names[t * 8 + k] = f"WG{t}: {nm}"
    for k, nm in enumerate(("item start", "X free seen", "next S issued", "dS part 0 seen -> dQ", "dS last part seen", "dQ issued", "next dP issued")):
        names[16 + t * 8 + k] = f"MMA (item of tile {t}): {nm}"
for t in (0, 1):
    names[32 + t] = f"MMA: dP MMAs of tile {t} accepted (this s)"
    names[34 + t] = f"MMA: commit y_full of tile {t} returned"
t0 = min(v for v in p if v > 0)
for v, nm in sorted((v - t0, f"s={8 + i // 64} {names.get(i % 64, i % 64)}") for i, v in enumerate(p) if v > 0):
    print(f"{v:7d}  {nm}")
