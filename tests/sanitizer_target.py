"""compute-sanitizer target (not a pytest file): one small call of every kernel family."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import flash_attention_metal_b200 as fa

n, d = 300, 64
for dd, nn, causal in ((64, 300, True), (128, 200, False)):
    q = torch.randn((2, nn, dd), device="cuda")
    o = torch.empty_like(q)
    for f in (fa.naive_attention, fa.flash_attention, fa.flash_attention_v2):
        f(q[0], q[0], q[0], o[0], nn, dd, dd ** -0.5, causal)
    fa.flash_attention_v2_batched(q, q, q, o, nn, dd, dd ** -0.5, 2 * nn * dd, nn * dd, causal, 1, 2)
    Q, K, V, dO = (torch.randn((2, nn, dd), device="cuda").to(torch.bfloat16) for _ in range(4))
    O = torch.empty_like(Q); L = torch.empty((2, nn), device="cuda")
    fa.flash_attention_v4_half(Q, K, V, O, nn, dd, dd ** -0.5, 2 * nn * dd, nn * dd, L, causal, 1, 2, fa.BF16)
    g = [torch.empty((2, nn, dd), device="cuda") for _ in range(3)]
    wsb = fa.workspace_bytes_backward(nn, dd, 1, 2); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    fa.flash_attention_backward(Q, K, V, O, dO, L, *g, nn, dd, dd ** -0.5, 2 * nn * dd, nn * dd, causal, 1, 2, fa.BF16, ws, wsb)
    Kr = torch.randn((2, 77, dd), device="cuda").to(torch.bfloat16)
    fa.flash_attention_v4_half_rect(Q, Kr, Kr, O, nn, 77, dd, dd ** -0.5, 2 * nn * dd, nn * dd, 2 * 77 * dd, 77 * dd, L, 1, 2, fa.BF16)
    torch.cuda.synchronize()
    assert torch.isfinite(O.float()).all() and all(torch.isfinite(x).all() for x in g)
print("sanitizer target ok")
