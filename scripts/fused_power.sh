#!/bin/bash
# GPU diagnostic: SM clock and board power while the fused draw + projection kernel runs (combined / no draw / no MMA)
for diag in 0 1 2; do
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 50 > gpurun_out/clk_$diag.csv &
  SMI=$!
  REPS=${REPS:-800} NWS=28 DIAGS=$diag python scripts/fused_bench.py
  kill $SMI
  python - <<PY
import numpy as np
rows=[l.strip().split(", ") for l in open("gpurun_out/clk_$diag.csv") if l.count(",")==2]
rows=[r for r in rows if float(r[1])>400]
if rows:
    print("  diag $diag: busy samples %d, SM clock median %.0f MHz (min %.0f), power median %.0f W (max %.0f), sw_power_cap active in %d" % (len(rows), np.median([float(r[0]) for r in rows]), min(float(r[0]) for r in rows), np.median([float(r[1]) for r in rows]), max(float(r[1]) for r in rows), sum(r[2].startswith("Active") for r in rows)))
PY
done
