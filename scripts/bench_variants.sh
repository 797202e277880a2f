#!/bin/bash
# GPU diagnostic: c5 step time and per-phase times for several tuning-knob settings (hx_ctx_set_tuning what=value)
for t in "$@"; do
  echo "== tune: $t"
  python bench.py --steps ${STEPS:-20} --warmup 3 --ess-iters 0 --no-e2e --no-cpu-baseline $t 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(round(d['ms_per_step'], 3), 'ms/step', d['phase_ms_per_half_move'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), d['acceptance_rate'])"
done
