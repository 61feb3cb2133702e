"""Development probe: timeline (SM cycles) of one forward CTA's hand-offs; needs a library built
with -DFA_FWD_TRACE (make EXTRA=-DFA_FWD_TRACE) selected through FA_B200_LIB."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa
L = fa.lib()
L.fa_debug_set_prof_buffer.argtypes = [ctypes.c_void_p]
B, H, n, d, causal = 1, 16, 16384, int(sys.argv[1]) if len(sys.argv) > 1 else 128, False
Q, K, V = (torch.randn((B, H, n, d), device="cuda").to(torch.bfloat16) for _ in range(3))
O = torch.empty_like(Q); Ls = torch.empty((B, H, n), device="cuda")
prof = torch.zeros(320, dtype=torch.int64, device="cuda")
for _ in range(int(os.environ.get("WARM", "2"))):
    fa.flash_attention_v4_half(Q, K, V, O, n, d, d ** -0.5, H * n * d, n * d, Ls, causal, B, H, fa.BF16)
L.fa_debug_set_prof_buffer(prof.data_ptr())
fa.flash_attention_v4_half(Q, K, V, O, n, d, d ** -0.5, H * n * d, n * d, Ls, causal, B, H, fa.BF16)
torch.cuda.synchronize()
L.fa_debug_set_prof_buffer(None)
p = prof.cpu().tolist()
if p[259] > p[257]:
    print(f"SM clock under load (last CTA, {p[259] - p[257]} ns): {(p[258] - p[256]) / (p[259] - p[257]) * 1000:.0f} MHz")
if p[263] > p[260]:
    print(f"CTA (0,0,0): {p[264]} iterations; entry -> first S ready {p[261] - p[260]} cycles, tile loop {p[262] - p[261]}, "
          f"epilogue (last P published -> all warps done) {p[263] - p[262]}")
p = p[:256]
names = {}
for t in (0, 1):
    for k, nm in enumerate(("S ready", "S in registers", "row max known", "P half 0 published", "P half 1 published")):
        names[t * 5 + k] = f"softmax{t}: {nm}"
names[18] = "MMA: loop top, waiting for V_j"
names[19] = "MMA: V_j landed, waiting for K_j+1"
names[20] = "MMA: K_j+1 landed"
names[40] = "TMA: stage free -> load K_j"
names[41] = "TMA: stage free -> load V_j"
for t in (0, 1):
    names[21 + t * 5] = f"MMA: tile{t} P half 0 seen -> issue PV"
    names[22 + t * 5] = f"MMA: tile{t} P half 1 seen -> issue PV"
    names[23 + t * 5] = f"MMA: tile{t} PV issued"
    names[25 + t * 5] = f"MMA: tile{t} next S issued"
t0 = min(v for v in p if v > 0)
ev = sorted((v - t0, f"j={8 + i // 64} {names.get(i % 64, i % 64)}") for i, v in enumerate(p) if v > 0)
for v, nm in ev:
    print(f"{v:7d}  {nm}")
