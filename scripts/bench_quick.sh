#!/bin/bash
# usage: bench_quick.sh [ENV=VAL ...]  -- one short bench line summary
env "$@" python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('us/iter %.1f  K1 %.1f us  e2e %.1f ms  it/s %.0f' % (d['us_per_lm_iteration'], d['roofline']['kernel_ms']*1e3, d['e2e']['ms_per_step'], d['lm_iters_per_sec']))"
