set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "tensorcore" 2>&1 | tail -15
N=4096 timeout 600 python scripts/diag_tc.py 2>&1 | grep -v "^$" | tail -40
for p in bf16x3 f16f8; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --ess-iters 0 --precision $p 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('RES', '$p', j['ms_per_step'], j['phase_ms_per_half_move'], j['clocks']['sm_mhz'], j['acceptance_rate'])"
done
