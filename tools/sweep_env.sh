#!/bin/bash
# Development sweep of the env step kernel's knobs.
out=gpurun_out/sweep3.txt; : > $out
run() {
  r=$(env "$@" python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-gae --no-rollout 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4e steps/s  kernel %.2f us  %.0f GB/s  frac %.3f' % (d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['achieved'], d['roofline']['frac']))" 2>&1 | tail -1)
  echo "$*  $r" | tee -a $out
}
run MSW_VEC_MODE=1
run MSW_VEC_MODE=2
for v in 0 1 2 3; do for blk in 32 64; do for bpw in 2 3 4; do
  run MSW_VARIANT=$v MSW_BLOCK=$blk MSW_BPW=$bpw
done; done; done
