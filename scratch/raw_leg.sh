#!/bin/bash
python bench.py --no-cpu --no-e2e > gpurun_out/raw_bench.json 2>gpurun_out/raw_bench.err; tail -3 gpurun_out/raw_bench.err
python -c "
import json; d=json.load(open('gpurun_out/raw_bench.json')); print(d['value'], d['roofline']['frac'])
for k,v in d['config']['raw_input'].items():
    if isinstance(v, dict): print(k, v['ms_per_step'], v['roofline_frac'], v['kernels_per_step'])"
