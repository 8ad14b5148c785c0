#!/bin/bash
# A/B of an experimental build (libslr_b200_exp.so) against the shipped library, same box
for V in base exp base exp; do
if [ $V = exp ]; then export SLR_B200_LIB=$PWD/structure-light-reconstructor_b200/libslr_b200_exp.so; else unset SLR_B200_LIB; fi
python bench.py --no-cpu --no-e2e --steps 30 > gpurun_out/ab.json 2>gpurun_out/ab.err
python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('$V', round(d['value']), round(d['roofline']['frac'],4), round(d['ms_per_step'],4), [round(v['roofline_frac'],4) for v in d['config']['variants']], {k: round(v['ms_per_step'],3) for k,v in d['config']['raw_input'].items() if isinstance(v, dict)})"
done
