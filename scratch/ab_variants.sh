#!/bin/bash
# A/B of candidate builds (scratch/variants/libslr_<v>.so) on one box: the bench's kernel-only legs (main, noisy, corrected,
# raw); a variant named with a trailing '+' runs the parity files first.
# usage (on the GPU box): scratch/ab_variants.sh v6 v4b v5b+ v6
mkdir -p gpurun_out
n=0
for A in "$@"; do
    V=${A%+}; n=$((n + 1))
    export SLR_B200_LIB=$PWD/scratch/variants/libslr_$V.so
    if [ "$A" != "$V" ]; then
        (timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_routes.py -m gpu -x -q 2>&1 | tail -4) > gpurun_out/ab_${V}_pytest.log
        echo "== $V parity: $(tail -1 gpurun_out/ab_${V}_pytest.log)"
    fi
    python bench.py --no-cpu --no-e2e --steps 20 > gpurun_out/ab_${V}_$n.json 2> gpurun_out/ab_${V}_$n.err
    python -c "
import json; d=json.load(open('gpurun_out/ab_${V}_$n.json')); print('$V', round(d['value']), round(d['roofline']['frac'],4), round(d['ms_per_step'],4), [round(v['ms_per_step'],4) for v in d['config']['variants']], {k: round(v['ms_per_step'],3) for k,v in d['config']['raw_input'].items() if isinstance(v, dict)})"
done
