#!/bin/bash
# GPU: bench (c5) + ncu launch list + ncu --set full of the fused projection kernel.  usage (under gpurun): bash scripts/r02_profile.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['phase_ms_per_half_move'], d['clocks'])"
B="python bench.py --steps 2 --warmup 1 --ess-iters 0 --no-e2e --no-cpu-baseline --clock-period 0"
$B > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_tc_proj_fused -s 2 -c 1 -o gpurun_out/${TAG}_fused_full -f $B > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/${TAG}_fused_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_fused_full_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_fused_full.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_fused_full_sass.csv 2>/dev/null
