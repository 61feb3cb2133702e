#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the attention hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the headline metric and the
north_star target are quoted on; it fits one GPU): causal attention, bf16, B=1,
H=16, N=16384, d=128, forward + backward.  One "step" = one forward + one backward
over that batch through the C ABI of libflash_attn_b200.so.  With N GPUs every rank
runs the same batch on its own GPU (batch x heads sharding, no data-path collective:
heads never interact, kernels.metal:622) -> weak scaling, value = whole-job TFLOP/s.

FLOP accounting: forward 4*B*H*N^2*d, halved when causal;
backward 2.5x forward; softmax flops not counted.

Printed keys beyond the base contract:
  roofline      the dominant kernel against the measured bf16 tensor peak
  cpu_baseline  the oracle (CPU port of the reference's verifier) on a bounded sample
  e2e           same metric through the host-buffer C-ABI call (H2D + kernels + D2H)
  fwd_tflops / bwd_tflops / fwd_ms / bwd_ms   the two passes separately
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, H, N, D = 1, 16, 16384, 128
CAUSAL = True
WORKLOAD = "causal attention fwd+bwd, bf16, B=1 H=16 N=16384 d=128 (BASELINE configs[2])"


def fwd_flops(b=B, h=H, n=N, d=D, causal=CAUSAL) -> float:
    return 4.0 * b * h * n * n * d * (0.5 if causal else 1.0)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "MEASURED_PEAKS.json"}
    # B200_PROFILING.md fallback
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for (_, l) in self.lines]
        sm, smax, reasons, power = [], [], set(), []
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path on host cores
# ---------------------------------------------------------------------------------
def cpu_forward_sample(n: int, heads: int, threads: int, prefer_ref: bool = True):
    """Causal forward over `heads` independent heads of length n, d=128, on `threads` host
    threads.  Uses oracle/_ref (the reference's loops compiled from main.mm:550-578) when
    present, else the oracle port.  Returns (seconds, flops, kind)."""
    import numpy as np

    import oracle

    scale = float(D ** -0.5)
    rng = np.random.default_rng(0)
    q, k, v = (rng.uniform(-1, 1, (heads, n, D)).astype(np.float32) for _ in range(3))
    o = np.empty_like(q)
    use_ref = prefer_ref and oracle.have_ref()
    if use_ref:
        R = oracle.ref()
        from concurrent.futures import ThreadPoolExecutor

        def one(hh):
            R.ref_forward_causal(q[hh], k[hh], v[hh], o[hh], n, D, scale)  # ctypes drops the GIL

        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, range(heads)))
        dt = time.perf_counter() - t0
    else:
        oracle.lib().oracle_set_num_threads(threads)
        t0 = time.perf_counter()
        for hh in range(heads):
            oracle.forward(q[hh], k[hh], v[hh], scale, True)
        dt = time.perf_counter() - t0
    return dt, fwd_flops(1, heads, n, D, True), ("reference" if use_ref else "port")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_s = 2048
    heads = cores
    for _ in range(args.warmup):
        cpu_forward_sample(512, heads, cores)
    times = []
    kind = "port"
    for _ in range(args.steps):
        dt, flops, kind = cpu_forward_sample(n_s, heads, cores)
        times.append(dt)
    total = sum(times)
    value = flops * args.steps / total / 1e12
    sample = (f"forward only, causal, fp32, {heads} independent heads x N={n_s} x d={D} per step "
              f"({'reference CPU loops main.mm:550-578 compiled into oracle/_ref' if kind == 'reference' else 'oracle port'}, "
              f"one head per host thread)")
    line = {
        "impl": "reference", "metric": "attention_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import flash_attention_metal_b200 as fa

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fa.lib()
    scale = float(D ** -0.5)
    st = torch.cuda.current_stream()

    g = torch.Generator(device="cuda").manual_seed(42 + rank)
    mk = lambda: torch.rand((B, H, N, D), device="cuda", generator=g).mul_(2).sub_(1).to(torch.bfloat16)
    Q, K, V, dO = mk(), mk(), mk(), mk()
    O = torch.empty_like(Q)
    L = torch.empty((B, H, N), device="cuda")
    dQ, dK, dV = (torch.empty((B, H, N, D), device="cuda") for _ in range(3))
    wsb = fa.workspace_bytes_backward(N, D, B, H)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")
    bs, hs = H * N * D, N * D

    def fwd():
        fa.flash_attention_v4_half(Q, K, V, O, N, D, scale, bs, hs, L, CAUSAL, B, H, fa.BF16, st)

    have_bwd = True

    def bwd():
        fa.flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, N, D, scale, bs, hs, CAUSAL, B, H, fa.BF16, ws, wsb, st)

    fwd()
    try:
        bwd()
    except fa.FlashAttnError as e:
        if "not implemented" not in str(e):
            raise
        have_bwd = False
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        fwd()
        if have_bwd:
            bwd()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    # K steps; an event after each pass so forward and backward are also known separately
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    fa.reset_launch_count()
    barrier()
    t0 = time.time()
    evs[0].record(st)
    for i in range(args.steps):
        fwd()
        evs[2 * i + 1].record(st)
        if have_bwd:
            bwd()
        evs[2 * i + 2].record(st)
    barrier()
    t1 = time.time()
    launches = fa.launch_count()
    clocks = sampler.stop(t0, t1)
    total_ms = evs[0].elapsed_time(evs[-1])
    fwd_ms = statistics.mean(evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(args.steps))
    bwd_ms = statistics.mean(evs[2 * i + 1].elapsed_time(evs[2 * i + 2]) for i in range(args.steps)) if have_bwd else None
    if world > 1:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    step_flops = fwd_flops() * (3.5 if have_bwd else 1.0)
    value = step_flops * world / (ms_per_step * 1e-3) / 1e12

    # ---- e2e: host buffers in, host buffers out, through the host-buffer C-ABI call ----
    e2e = None
    if not args.no_e2e:
        hq, hk, hv = (torch.empty((B, H, N, D), dtype=torch.bfloat16).pin_memory() for _ in range(3))
        for hsrc, dsrc in ((hq, Q), (hk, K), (hv, V)):
            hsrc.copy_(dsrc)
        ho = torch.empty((B, H, N, D), dtype=torch.bfloat16).pin_memory()
        hl = torch.empty((B, H, N), dtype=torch.float32).pin_memory()
        call = lambda: fa.host_attention_half(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), hl.data_ptr(),
                                              N, D, scale, CAUSAL, B, H, fa.BF16)
        for _ in range(2):
            call()
        barrier()
        w0 = time.perf_counter()
        esteps = max(3, min(args.steps, 10))
        for _ in range(esteps):
            call()  # synchronous: returns after the D2H copy of O and L
        barrier()
        e_ms = (time.perf_counter() - w0) * 1e3 / esteps
        if world > 1:
            t = torch.tensor([e_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        nbytes = B * H * N * D * 2
        e2e = {"value": fwd_flops() * world / (e_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": 3 * nbytes, "d2h_bytes_per_step": nbytes + B * H * N * 4,
               "what": "forward only through fa_host_attention_half (pinned host Q,K,V -> device, kernel, O and L -> host)"}
        assert torch.equal(ho.cuda(), O), "e2e result differs from the device-resident result"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    fwd_tf = fwd_flops() / (fwd_ms * 1e-3) / 1e12
    bwd_tf = 2.5 * fwd_flops() / (bwd_ms * 1e-3) / 1e12 if have_bwd else None
    roof = {"bound": "tensor", "kernel": "fwd_tc_kernel<128,bf16> (flash_attention_v4_half)", "achieved": fwd_tf,
            "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": fwd_tf / peaks["bf16_tflops"],
            "peak_source": peaks["source"] + " (cuBLAS bf16 burst)", "frac_of_nominal_2250": fwd_tf / 2250.0,
            "frac_of_sustained": fwd_tf / peaks["bf16_tflops_sustained"] if peaks["bf16_tflops_sustained"] else None,
            "traffic": None, "flops_per_launch": fwd_flops(), "launch_ms": fwd_ms}
    prof = os.path.join(ROOT, "profiles", "fwd_traffic.json")
    if os.path.exists(prof):
        roof["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")

    cpu = None
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        # bounded sample: ~10-30 s of CPU work on the host cores
        n_s = 4096
        dt, flops, kind = cpu_forward_sample(n_s, cores, cores, prefer_ref=False)
        cpu = {"value": flops / dt / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": kind, "seconds": dt,
               "sample": f"forward only, causal, fp32 oracle port (OpenMP over rows), {cores} heads x N={n_s} x d={D}"}

    line = {
        "metric": "attention_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pass": "fwd+bwd" if have_bwd else "fwd only (backward not built)",
                   "B": B, "H_per_gpu": H, "N": N, "d": D, "causal": CAUSAL, "sharding": f"batch x heads, {world} rank(s), no collective",
                   "l2": "inputs (Q,K,V,dO = 256 MiB) exceed the 126 MB L2; no explicit flush"},
        "fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "fwd_tflops": fwd_tf, "bwd_tflops": bwd_tf,
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
