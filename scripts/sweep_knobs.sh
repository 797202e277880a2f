# usage: bash scripts/sweep_knobs.sh "<bench args 1>" "<bench args 2>" ...   (one bench.py run per argument string)
for t in "$@"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --ess-iters 0 $t 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('RES', sys.argv[1], '| ms/step', round(j['ms_per_step'],3), j['phase_ms_per_half_move'], 'clk', j['clocks']['sm_mhz'], 'acc', round(j['acceptance_rate'],5))" "$t"
done
