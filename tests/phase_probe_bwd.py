"""Development probe: per-phase cycle counts of one backward dK/dV CTA."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa
L = fa.lib()
L.fa_debug_set_prof_buffer.argtypes = [ctypes.c_void_p]
for (B, H, n, d, causal) in [(1, 16, 16384, 128, False), (1, 16, 16384, 128, True)]:
    Q, K, V, dO = (torch.randn((B, H, n, d), device="cuda").to(torch.bfloat16) for _ in range(4))
    O = torch.empty_like(Q); Ls = torch.empty((B, H, n), device="cuda")
    dQ, dK, dV = (torch.empty((B, H, n, d), device="cuda") for _ in range(3))
    wsb = fa.workspace_bytes_backward(n, d, B, H); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, n, d, d ** -0.5, H * n * d, n * d, Ls, causal, B, H, fa.BF16)
    run = lambda: fa.flash_attention_backward(Q, K, V, O, dO, Ls, dQ, dK, dV, n, d, d ** -0.5, H * n * d, n * d, causal, B, H, fa.BF16, ws, wsb)
    run(); run()
    prof = torch.zeros(32, dtype=torch.int64, device="cuda")
    L.fa_debug_set_prof_buffer(prof.data_ptr()); run(); torch.cuda.synchronize(); L.fa_debug_set_prof_buffer(None)
    p = prof.cpu().tolist()
    for wg in (0, 1):
        nt, wx, p1, wy, p2, tot, top, top2 = p[8 * wg:8 * wg + 8]
        print(f"dkdv N={n} causal={int(causal)} WG{wg}: tiles {nt}  wait X {wx/nt:.0f}  phase1 {p1/nt:.0f}  wait Y {wy/nt:.0f}  phase2 {p2/nt:.0f}  total/tile {tot/nt:.0f}  [stat store {top/nt:.0f}, fetch+barrier {top2/nt:.0f}]")
    nt = p[0]
    print(f"   MMA thread: wait P {p[16]/nt:.0f}  wait dS {p[17]/nt:.0f}  loop/tile {p[18]/nt:.0f}  (MMA work per tile 2048 cycles)")
