#!/bin/bash
for V in 0 1 0 1; do
if [ $V = 1 ]; then export SLR_FLOW_NO_PREFETCH=1; else unset SLR_FLOW_NO_PREFETCH; fi
python bench.py --no-cpu --no-e2e --no-variants --steps 30 > gpurun_out/ab.json 2>gpurun_out/ab.err
python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('noprefetch=$V', round(d['value']), round(d['roofline']['frac'],4), round(d['ms_per_step'],4))"
done
