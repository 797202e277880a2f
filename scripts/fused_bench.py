"""GPU diagnostic: the fused draw + projection kernel (hx_tc_microbench mode 3) against the staged draw / GEMM,
over RNG warp counts, rotation masks and the no-draw / no-MMA diagnostics."""
import sys, os, ctypes as C
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "h-emcee_b200")]
import hemcee_b200 as hx
from hemcee_b200 import _lib, _rt
if os.environ.get("HX_LIB"):
    _lib.LIB_PATH = os.environ["HX_LIB"]      # diagnostic builds (e.g. other pipeline depths)
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
lib = _lib.load(); ctx = _rt.ctx(dev)
n = int(os.environ.get("N", 32768)); rows = int(os.environ.get("ROWS", n)); d = int(os.environ.get("D", 1000))
reps = int(os.environ.get("REPS", 20))
lib.hx_ctx_set_precision(ctx, hx.config.PRECISIONS["bf16x3"])
def run(mode, label):
    ms = (C.c_double * 3)()
    st = _rt.stream(dev)
    _lib.check(lib.hx_tc_microbench(ctx, mode, n, rows, d, 2, ms, st))
    _lib.check(lib.hx_tc_microbench(ctx, mode, n, rows, d, reps, ms, st))
    print(f"{label:50s} {ms[0] / reps:8.3f} ms/rep", flush=True)
NWS = [int(v) for v in os.environ.get("NWS", "28,24,20,16").split(",")]
FMTS = [int(v) for v in os.environ.get("FMTS", "1,2").split(",")]      # 1 bf16 hi/lo, 2 fp16 + fp8 (hx_ctx_set_tuning 21)
MASKS = [int(v, 0) for v in os.environ.get("SLICES", "0").split(",")]
DIAGS = [int(v) for v in os.environ.get("DIAGS", "0,1,2").split(",")]
NAMES = {0: "", 1: " no draw (MMA + TMA only)", 2: " no MMA (draw + TMA only)", 3: " draw arithmetic only", 4: " draw + local stores only"}
for fmt in FMTS:
  for nw in NWS:
    for mask in MASKS:
        for diag in DIAGS:
            lib.hx_ctx_set_tuning(ctx, 21, fmt)
            lib.hx_ctx_set_tuning(ctx, 16, nw); lib.hx_ctx_set_tuning(ctx, 17, mask); lib.hx_ctx_set_tuning(ctx, 19, diag)
            run(3, f"fused fmt={fmt} nw={nw} slices={mask}{NAMES.get(diag, ' diag %d' % diag)}")
lib.hx_ctx_set_tuning(ctx, 19, 0)
if os.environ.get("STAGED", "0") == "1":
    run(0, "staged: draw alone"); run(1, "staged: GEMM alone"); run(2, "staged: both")
