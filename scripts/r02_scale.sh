#!/bin/bash
# the driver's scaling invocation for one N (run under gpurun --gpus N): full bench line incl. e2e and ESS -> gpurun_out/r02_scale_nN.json
N=$1
if [ "$N" = "1" ]; then python bench.py --gpus 1 --steps 10 --warmup 3 2>gpurun_out/r02_scale_n1.err | tail -1 > gpurun_out/r02_scale_n1.json
else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/r02_scale_n$N.err | tail -1 > gpurun_out/r02_scale_n$N.json; fi
python - <<PY
import json
d = json.load(open("gpurun_out/r02_scale_n$N.json"))
print("N=$N", round(d["ms_per_step"], 3), "ms/step value %.4g e2e %.4g ess %.4g" % (d["value"], d["e2e"]["value"], d["ess"]["value"]), d["checksum"]["ensemble"], d["acceptance_rate"], d["phase_ms_per_half_move"], d["clocks"].get("sm_mhz"), d["ess"].get("var_over_target_minmax"))
PY
