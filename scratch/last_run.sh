#!/bin/bash
# bench line + one ncu --set full capture of the shipped build, then the GPU suite and smoke()
tag=$1; mkdir -p gpurun_out
timeout 120 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); print('value', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], [round(v['ms_per_step'],4) for v in d['config']['variants']], {k: round(v['ms_per_step'],3) for k,v in d['config']['raw_input'].items() if isinstance(v, dict)})"
timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_fused_flow -s 3 -c 1 -f -o gpurun_out/${tag}_flow \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-variants > gpurun_out/${tag}_ncu_full.log 2>&1
timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/${tag}_smoke.log
