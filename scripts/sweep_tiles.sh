#!/bin/bash
# sweep the tiles-per-CTA knob of the hardware-scheduled kernels (0 = persistent kernel)
mkdir -p gpurun_out
OUT=gpurun_out/sweep_tiles.log
: > $OUT
for cfg in "0,0,0" "1,1,1" "2,2,2" "4,4,4" "8,8,8" "16,16,16" "32,32,32"; do
	ELB_TILES_PER_CTA=$cfg timeout 200 python bench.py --skip-e2e --skip-cpu --steps 30 --warmup 3 2>/dev/null | \
		python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['roofline']['all_kernels']
print('$cfg', 'value', d['value'], 'K1', k['K1_fill_pattern']['achieved'], 'K2', k['K2_verify_pattern']['achieved'], 'K3', k['K3_fill_random_pct100']['achieved'])" >> $OUT
done
cat $OUT
