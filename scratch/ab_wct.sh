#!/bin/bash
for V in 0 1 0 1; do
if [ $V = 1 ]; then export SLR_FLOW_GENERIC_W=1; else unset SLR_FLOW_GENERIC_W; fi
python bench.py --no-cpu --no-e2e --steps 30 > gpurun_out/ab.json 2>gpurun_out/ab.err
python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('generic=$V', round(d['value']), round(d['roofline']['frac'],4), round(d['ms_per_step'],4), [round(v['roofline_frac'],4) for v in d['config']['variants']], {k: round(v['ms_per_step'],3) for k,v in d['config']['raw_input'].items() if isinstance(v, dict)})"
done
