"""Development probe: where does the single-process ring stall?  Prints progress to stderr; polls instead of blocking."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa

mech = int(sys.argv[1]) if len(sys.argv) > 1 else -1
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = fa.lib()
L.fa_debug_ring_flags.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
L.fa_debug_mgpu_ring.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.fa_debug_mgpu_ring.restype = ctypes.c_void_p
L.fa_debug_set_ring_write_mechanism.argtypes = [ctypes.c_int]
say = lambda *a: print(*a, file=sys.stderr, flush=True)
P, H, D, n_local = 2, 2, 128, 512
scale = D ** -0.5
if mech >= 0:
    L.fa_debug_set_ring_write_mechanism(mech)
mk = lambda i: torch.randn((H, n_local, D), device=f"cuda:{i}").to(torch.bfloat16)
Q, K, V = ([mk(i) for i in range(P)] for _ in range(3))
O = [torch.zeros_like(q) for q in Q]
Ls = [torch.zeros((H, n_local), device=f"cuda:{i}") for i in range(P)]
if warm:  # run the forward kernel once on every device first (module loading out of the way)
    for i in range(P):
        torch.cuda.set_device(i)
        fa.flash_attention_v4_half(Q[i], K[i], V[i], O[i], n_local, D, scale, H * n_local * D, n_local * D, Ls[i], False, 1, H, fa.BF16)
        torch.cuda.synchronize(i)
    say("warmed up")
for i in range(P):
    torch.cuda.synchronize(i)
grp = fa.Mgpu(list(range(P)))
say("group created")
t0 = time.time()
grp.ring_forward(Q, K, V, O, Ls, n_local, D, H, scale, False, fa.BF16)
say(f"enqueue returned after {time.time() - t0:.3f} s; write mechanism now {L.fa_debug_ring_write_mechanism()}")
streams = [torch.cuda.ExternalStream(grp.stream(i), device=f"cuda:{i}") for i in range(P)]
for k in range(30):
    done = [s.query() for s in streams]
    if all(done):
        break
    time.sleep(0.1)
say("streams done:", done)
buf = (ctypes.c_uint32 * 40)()
for i in range(P):
    rc = L.fa_debug_ring_flags(L.fa_debug_mgpu_ring(grp.handle, i), buf, 40)
    say(f"rank {i} flags rc={rc}: kv_ready={list(buf[0:4])} kv_done={list(buf[16:20])} acc={list(buf[32:34])}")
if all(done):
    print("MGPU DEBUG OK", flush=True)
os._exit(0 if all(done) else 3)   # never block in destructors when stalled
