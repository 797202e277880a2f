"""GPU diagnostic: draw (ALU) vs projection GEMM (tensor) overlap and power (hx_tc_microbench)."""
import sys, os, subprocess, threading, time, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h-emcee_b200")]
import hemcee_b200 as hx
from hemcee_b200 import _lib, _rt
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
lib = _lib.load(); ctx = _rt.ctx(dev)
n = int(os.environ.get("N", 32768)); rows = int(os.environ.get("ROWS", n)); d = 1000
reps = int(os.environ.get("REPS", 200))
user = torch.cuda.Stream()
def sample(stop, out):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"], stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        line = p.stdout.readline()
        if line: out.append(line.strip().split(", "))
    p.terminate()
PRECS = os.environ.get("PRECS", "bf16x3,f16f8").split(",")
CPS = [int(c) for c in os.environ.get("CPS", "2,3,4,6").split(",")]
runs = []
PAIRS = [int(c) for c in os.environ.get("PAIRS", "0,1").split(",")]
for pr in PRECS:
    runs += [(pr, 0, 0, 4, f"{pr}: draw alone 4")]
    for pair in PAIRS:
        runs += [(pr, pair, 1, 4, f"{pr} pair={pair}: gemm alone")] + [(pr, pair, 2, c, f"{pr} pair={pair}: both, draw {c} CTA/SM") for c in CPS]
NOSTORE = int(os.environ.get("NOSTORE", 0))
lib.hx_ctx_set_tuning(ctx, 10, NOSTORE)
for pr, pair, mode, cps, name in runs:
    lib.hx_ctx_set_precision(ctx, hx.config.PRECISIONS[pr])
    lib.hx_ctx_set_tuning(ctx, 8, 1 if pair else 0)
    lib.hx_ctx_set_tuning(ctx, 9, 1 if pair == 2 else 0)        # pair = 2: CTA-pair kernel launched wave by wave
    lib.hx_ctx_set_tuning(ctx, 0, cps)
    ms = (C.c_double * 3)()
    with torch.cuda.stream(user):
        st = _rt.stream(dev)
        _lib.check(lib.hx_tc_microbench(ctx, mode, n, rows, d, 2, ms, st))   # warm
        stop, out = threading.Event(), []
        th = threading.Thread(target=sample, args=(stop, out)); th.start(); time.sleep(0.2)
        _lib.check(lib.hx_tc_microbench(ctx, mode, n, rows, d, reps, ms, st))
    stop.set(); th.join()
    rows_ = [r for r in out if len(r) == 3]
    busy = rows_[len(rows_) // 3:]          # later samples: power reading has settled
    clk = np.median([float(r[0]) for r in busy]) if busy else float("nan")
    pw = np.max([float(r[1]) for r in busy]) if busy else float("nan")
    cap = sum(r[2].lower().startswith("active") for r in busy)
    print(f"{name:40s} total {ms[0] / reps:7.3f} ms/rep  gemm-stream {ms[1] / reps:7.3f}  draw-stream {ms[2] / reps:7.3f}   sm clock {clk:.0f} MHz  max power {pw:.0f} W  power-cap {cap}/{len(busy)}", flush=True)
    time.sleep(0.5)
