"""GPU diagnostic: fused projection alone (hx_tc_microbench mode 3, c5 shape) over values of one tuning knob: KNOB=24 VALUES=0,800,1600"""
import sys, os, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h-emcee_b200")]
import hemcee_b200 as hx
from hemcee_b200 import _lib, _rt
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
lib = _lib.load(); ctx = _rt.ctx(dev)
n = int(os.environ.get("N", 32768)); rows = int(os.environ.get("ROWS", n)); d = int(os.environ.get("D", 1000)); reps = int(os.environ.get("REPS", 20))
knob = int(os.environ.get("KNOB", 24)); vals = [int(v) for v in os.environ.get("VALUES", "0").split(",")]
lib.hx_ctx_set_precision(ctx, hx.config.PRECISIONS["auto"])
for v in vals + vals[:1]:
    lib.hx_ctx_set_tuning(ctx, knob, v)
    ms = (C.c_double * 3)(); st = _rt.stream(dev)
    _lib.check(lib.hx_tc_microbench(ctx, 3, n, rows, d, 2, ms, st))
    _lib.check(lib.hx_tc_microbench(ctx, 3, n, rows, d, reps, ms, st))
    print(f"knob {knob} = {v}: {ms[0] / reps:8.3f} ms/rep", flush=True)
