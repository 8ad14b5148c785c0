#!/bin/bash
for V in 0 1 0 1; do
if [ $V = 1 ]; then export SLR_FUSED_GENERIC_W=1; else unset SLR_FUSED_GENERIC_W; fi
python bench.py --config 4 --batch 16 --no-cpu --no-e2e --steps 10 > gpurun_out/cfg4.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/cfg4.json')); print('generic=$V', round(d['value']), round(d['roofline']['frac'],4), round(d['ms_per_step'],4))"
done
