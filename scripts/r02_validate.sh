#!/bin/bash
# GPU validation of one build: gpu tests, smoke, bench (c5), ncu launch list, ncu --set full of the fused projection kernel.
# usage (under gpurun): bash scripts/r02_validate.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/${TAG}_tests.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${TAG}_bench.json
B="python bench.py --steps 2 --warmup 1 --ess-iters 0 --no-e2e --no-cpu-baseline --clock-period 0"
$B > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_tc_proj_fused -s 2 -c 1 -o gpurun_out/${TAG}_fused_full -f $B > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/${TAG}_fused_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_fused_full_raw.csv 2>/dev/null
