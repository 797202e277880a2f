set -x
python -m pytest tests -m gpu -x -q -k "tensorcore" 2>&1 | tail -3
for t in "" "--tune 7=1" "--tune 6=1" "--tune 6=2" "--tune 6=3" "--tune 6=5" "--tune 6=99"; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --ess-iters 0 $t 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('RES', '$t', j['ms_per_step'], j['phase_ms_per_half_move'], j['clocks']['sm_mhz'])"
done
