#!/bin/bash
for B in 4 16; do for F in 1 0; do
SLR_FUSED_FLOW=$F timeout 300 python bench.py --config 4 --batch $B --no-cpu --no-e2e --steps 10 > gpurun_out/cfg4.json 2>gpurun_out/cfg4.err
python -c "
import json; d=json.load(open('gpurun_out/cfg4.json')); print('B=$B flow=$F', round(d['value']), round(d['roofline']['frac'],4), round(d['ms_per_step'],4))"
done; done
