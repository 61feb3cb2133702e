#!/usr/bin/env python3
"""Second plot for the harness CSVs: absolute throughput instead of
the reference's speed-up-over-naive (plot_results.py:36-39).

    python harness/plot_tflops.py [benchmark_results*.csv ...] [-o tflops_plot.svg]
                                  [--tensor-peak 1641.3] [--fp32-peak 74] [--hbm-peak 6550.7]

For every CSV (one curve family per file: fp16 / bf16, causal / not) it draws, against N on a log axis,
  left panel : TFLOP/s of the fp32 kernel (FlashV2) and of the tensor-core kernel (FlashV4), log scale,
               with the fp32 FFMA peak and the measured bf16 tensor peak as horizontal lines;
  right panel: the tensor-core kernel as % of the tensor peak and, for the small-N memory-bound regime,
               its achieved HBM GB/s as % of the HBM peak.
FLOPs are recomputed from the ms columns (4 N^2 d, halved for causal files) exactly as the harness
computes them, so a CSV without the extra columns (the reference's own 10 columns) still plots.
Only the standard library is used, like plot_results.py, and the output is a standalone SVG."""
import argparse
import math
import os
import sys

COLORS = ["#1f77b4", "#d62728", "#2ca02c", "#9467bd", "#ff7f0e", "#8c564b", "#17becf", "#7f7f7f"]


def read_csv(path, d):
    """rows of (N, v2_tflops, v4_tflops, v4_gbs); parsing rules as plot_results.py:16-34"""
    causal = "causal" in os.path.basename(path)
    rows, started = [], False
    for line in open(path):
        line = line.strip()
        if "N,Naive(ms)" in line:
            started = True
            continue
        if not started or not line:
            continue
        f = line.split(",")
        if len(f) < 6:
            continue
        try:
            n, t2, t4 = int(f[0]), float(f[3]), float(f[5])
        except ValueError:
            continue
        flop = 4.0 * n * n * d * (0.5 if causal else 1.0)
        gbs = (4.0 * n * d * 2 + 4.0 * n) / (t4 * 1e6) if t4 > 0 else 0.0
        rows.append((n, flop / (t2 * 1e9) if t2 > 0 else 0.0, flop / (t4 * 1e9) if t4 > 0 else 0.0, gbs))
    return rows


class Panel:
    def __init__(self, x0, y0, w, h, xmin, xmax, ymin, ymax, logy):
        self.x0, self.y0, self.w, self.h = x0, y0, w, h
        self.xmin, self.xmax, self.ymin, self.ymax, self.logy = math.log2(xmin), math.log2(xmax), ymin, ymax, logy

    def x(self, n):
        return self.x0 + (math.log2(n) - self.xmin) / max(self.xmax - self.xmin, 1e-9) * self.w

    def y(self, v):
        if self.logy:
            v, lo, hi = math.log10(max(v, self.ymin)), math.log10(self.ymin), math.log10(self.ymax)
        else:
            lo, hi = self.ymin, self.ymax
        return self.y0 + self.h - (v - lo) / (hi - lo) * self.h

    def frame(self, out, title, ylabel, ns, yticks):
        out.append(f'<rect x="{self.x0}" y="{self.y0}" width="{self.w}" height="{self.h}" fill="none" stroke="#333"/>')
        out.append(f'<text x="{self.x0 + self.w / 2}" y="{self.y0 - 12}" text-anchor="middle" font-size="15">{title}</text>')
        out.append(f'<text x="{self.x0 - 46}" y="{self.y0 + self.h / 2}" text-anchor="middle" font-size="12" '
                   f'transform="rotate(-90 {self.x0 - 46} {self.y0 + self.h / 2})">{ylabel}</text>')
        for n in ns:
            out.append(f'<line x1="{self.x(n):.1f}" y1="{self.y0 + self.h}" x2="{self.x(n):.1f}" y2="{self.y0 + self.h + 5}" stroke="#333"/>')
            out.append(f'<text x="{self.x(n):.1f}" y="{self.y0 + self.h + 18}" text-anchor="middle" font-size="11">{n}</text>')
        for v in yticks:
            out.append(f'<line x1="{self.x0}" y1="{self.y(v):.1f}" x2="{self.x0 + self.w}" y2="{self.y(v):.1f}" stroke="#ddd"/>')
            out.append(f'<text x="{self.x0 - 6}" y="{self.y(v) + 4:.1f}" text-anchor="end" font-size="11">{v:g}</text>')
        out.append(f'<text x="{self.x0 + self.w / 2}" y="{self.y0 + self.h + 36}" text-anchor="middle" font-size="12">N (sequence length)</text>')

    def curve(self, out, pts, color, dash=""):
        pts = [(n, v) for n, v in pts if v > 0]
        if not pts:
            return
        d = " ".join(f"{self.x(n):.1f},{self.y(v):.1f}" for n, v in pts)
        extra = f' stroke-dasharray="{dash}"' if dash else ""
        out.append(f'<polyline points="{d}" fill="none" stroke="{color}" stroke-width="2"{extra}/>')
        for n, v in pts:
            out.append(f'<circle cx="{self.x(n):.1f}" cy="{self.y(v):.1f}" r="3" fill="{color}"/>')

    def hline(self, out, v, label):
        out.append(f'<line x1="{self.x0}" y1="{self.y(v):.1f}" x2="{self.x0 + self.w}" y2="{self.y(v):.1f}" stroke="#000" stroke-dasharray="6,4"/>')
        out.append(f'<text x="{self.x0 + self.w - 4}" y="{self.y(v) - 4:.1f}" text-anchor="end" font-size="11">{label}</text>')


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("csv", nargs="*", default=["benchmark_results.csv"])
    ap.add_argument("-o", "--output", default="tflops_plot.svg")
    ap.add_argument("--d", type=int, default=64, help="head dimension the CSV was taken with")
    ap.add_argument("--tensor-peak", type=float, default=1641.3, help="measured bf16 tensor peak, TFLOP/s")
    ap.add_argument("--fp32-peak", type=float, default=74.0, help="fp32 FFMA peak, TFLOP/s (nominal)")
    ap.add_argument("--hbm-peak", type=float, default=6550.7, help="measured HBM copy bandwidth, GB/s")
    a = ap.parse_args(argv)
    series = [(os.path.splitext(os.path.basename(p))[0].replace("benchmark_results", "").strip("_") or "fp16", read_csv(p, a.d))
              for p in a.csv if os.path.exists(p)]
    series = [(name, rows) for name, rows in series if rows]
    if not series:
        print("no rows found in", a.csv)
        return 1
    ns = sorted({r[0] for _, rows in series for r in rows})
    W, H = 1180, 560
    out = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" font-family="sans-serif">',
           f'<rect width="{W}" height="{H}" fill="white"/>']
    left = Panel(80, 50, 470, 400, ns[0], ns[-1], 0.01, 4000.0, True)
    right = Panel(660, 50, 470, 400, ns[0], ns[-1], 0.0, 100.0, False)
    left.frame(out, "attention throughput, single head (TFLOP/s, log scale)", "TFLOP/s", ns, [0.01, 0.1, 1, 10, 100, 1000])
    right.frame(out, "fraction of the B200 roofline (%)", "% of peak", ns, [0, 20, 40, 60, 80, 100])
    left.hline(out, a.tensor_peak, f"bf16 tensor peak {a.tensor_peak:g} (measured)")
    left.hline(out, a.fp32_peak, f"fp32 FFMA peak {a.fp32_peak:g}")
    ly = 470
    for i, (name, rows) in enumerate(series):
        c = COLORS[i % len(COLORS)]
        left.curve(out, [(r[0], r[2]) for r in rows], c)
        left.curve(out, [(r[0], r[1]) for r in rows], c, "5,4")
        right.curve(out, [(r[0], 100.0 * r[2] / a.tensor_peak) for r in rows], c)
        right.curve(out, [(r[0], 100.0 * r[3] / a.hbm_peak) for r in rows], c, "2,3")
        out.append(f'<rect x="{90 + i * 260}" y="{ly + 32}" width="14" height="4" fill="{c}"/>')
        out.append(f'<text x="{110 + i * 260}" y="{ly + 38}" font-size="12">{name}</text>')
    out.append(f'<text x="80" y="{ly + 62}" font-size="12">left: solid = FlashV4 (tcgen05 tensor cores), dashed = FlashV2 (fp32 FFMA).  '
               f'right: solid = FlashV4 % of tensor peak, dotted = FlashV4 achieved HBM GB/s as % of {a.hbm_peak:g} GB/s (small-N regime).</text>')
    out.append("</svg>")
    open(a.output, "w").write("\n".join(out))
    print("wrote", a.output, "from", ", ".join(n for n, _ in series))
    return 0


if __name__ == "__main__":
    sys.exit(main())
