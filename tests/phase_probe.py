"""Development probe: per-phase cycle counts of one forward CTA (softmax wait/work, MMA waits)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa
L = fa.lib()
L.fa_debug_set_prof_buffer.argtypes = [ctypes.c_void_p]
for (B, H, n, d, causal) in [(1, 16, 16384, 128, False), (1, 16, 16384, 128, True), (8, 12, 4096, 64, True)]:
    Q, K, V = (torch.randn((B, H, n, d), device="cuda").to(torch.bfloat16) for _ in range(3))
    O = torch.empty_like(Q); Ls = torch.empty((B, H, n), device="cuda")
    prof = torch.zeros(16, dtype=torch.int64, device="cuda")
    for _ in range(2):
        fa.flash_attention_v4_half(Q, K, V, O, n, d, d ** -0.5, H * n * d, n * d, Ls, causal, B, H, fa.BF16)
    L.fa_debug_set_prof_buffer(prof.data_ptr())
    fa.flash_attention_v4_half(Q, K, V, O, n, d, d ** -0.5, H * n * d, n * d, Ls, causal, B, H, fa.BF16)
    torch.cuda.synchronize()
    L.fa_debug_set_prof_buffer(None)
    p = prof.cpu().tolist()
    for t in (0, 1):
        nt, w, k, tot = p[4 * t:4 * t + 4]
        if nt:
            print(f"N={n} d={d} causal={int(causal)} softmax WG{t}: tiles {nt}  wait/tile {w/nt:.0f}  work/tile {k/nt:.0f}  total/tile {tot/nt:.0f} cycles")
    nt = max(p[0], p[4])
    print(f"   MMA thread: wait P per tile-iter {p[8]/nt:.0f}, wait KV {p[9]/nt:.0f}, loop total per iter {p[10]/nt:.0f} (MMA work per iter = {2*(2*128*128*d*2)//8192} tensor cycles)")
