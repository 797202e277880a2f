#!/bin/bash
# GPU diagnostic: c5 step time, phases and checksums at N ranks (run under gpurun --gpus N)
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 --ess-iters 0 --no-e2e "$@" 2>&1 | tail -1 > gpurun_out/bench_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/bench_n$N.json"))
print("N=$N", round(d["ms_per_step"], 3), "ms/step", d["phase_ms_per_half_move"], d["checksum"]["ensemble"], d["acceptance_rate"], "enqueue", round(d["host_enqueue_ms_per_step"], 3), d["clocks"].get("sm_mhz"))
PY
