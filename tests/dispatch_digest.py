"""Helper for test_gpu_dispatch.py (not a pytest file): run one fixed causal forward + backward and
print a digest of every output.  The dispatch order (argv[2] = L2 budget of a head group in MB, set
through the library's debug hook) must not change a single bit."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa

B, H, n, d = 2, 5, 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 64
if len(sys.argv) > 2:
    import ctypes
    fa.lib().fa_debug_set_l2_group_mb.argtypes = [ctypes.c_int]
    fa.lib().fa_debug_set_l2_group_mb(int(sys.argv[2]))
g = torch.Generator(device="cuda").manual_seed(3)
Q, K, V, dO = (torch.rand((B, H, n, d), device="cuda", generator=g).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(4))
O = torch.empty_like(Q); L = torch.empty((B, H, n), device="cuda")
scale = d ** -0.5
fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, True, B, H, fa.BF16)
dQ, dK, dV = (torch.empty((B, H, n, d), device="cuda") for _ in range(3))
wsb = fa.workspace_bytes_backward(n, d, B, H); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
fa.flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, n, d, scale, H * n * d, n * d, True, B, H, fa.BF16, ws, wsb)
torch.cuda.synchronize()
h = hashlib.sha256()
for t in (O, L, dQ, dK, dV):
    h.update(t.cpu().numpy().tobytes() if t.dtype != torch.bfloat16 else t.view(torch.int16).cpu().numpy().tobytes())
print("digest", h.hexdigest())
