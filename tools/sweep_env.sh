#!/bin/bash
# Development sweep of the env step kernel's launch knobs.
out=gpurun_out/sweep2.txt; : > $out
run() {
  r=$(env "$@" python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4e steps/s  kernel %.2f us  %.0f GB/s  frac %.3f' % (d['value'], d['roofline']['kernel_ms']*1e3, d['roofline']['achieved'], d['roofline']['frac']))" 2>&1 | tail -1)
  echo "$*  $r" | tee -a $out
}
for v in 0 2; do for blk in 32 64 128 256; do for bpw in 1 2 3 4 6; do
  run MSW_VARIANT=$v MSW_BLOCK=$blk MSW_BPW=$bpw
done; done; done
