#!/bin/bash
# One GPU call at the end of a round: parity, smoke, bench line, kernel table, launch list, one ncu --set full capture.
# usage (on the GPU box): scratch/final_run.sh <tag> [fallback-lib]   -> gpurun_out/<tag>_*
# If the in-tree library fails the GPU suite and a fallback library is given, the remaining steps run on the fallback
# (gpurun_out/<tag>_lib.txt says which library the numbers belong to).
tag=$1; fb=$2
mkdir -p gpurun_out
echo intree > gpurun_out/${tag}_lib.txt
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
rc=$?
tail -3 gpurun_out/${tag}_pytest.log
if [ $rc -ne 0 ] && [ -n "$fb" ]; then
    export SLR_B200_LIB=$PWD/$fb
    echo "fallback $fb" > gpurun_out/${tag}_lib.txt
    timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_fallback.log 2>&1
    tail -3 gpurun_out/${tag}_pytest_fallback.log
fi
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.log
# the fallback as a same-box reference for the headline kernel (kept only if it is more than 0.5 % faster)
if [ -n "$fb" ] && [ -z "$SLR_B200_LIB" ]; then
    ms() { python -c "import json,sys; print(json.load(open(sys.argv[1]))['ms_per_step'])" $1; }
    timeout 100 python bench.py --no-cpu --no-e2e --no-variants --steps 20 > gpurun_out/${tag}_ab_intree.json 2>/dev/null
    SLR_B200_LIB=$PWD/$fb timeout 100 python bench.py --no-cpu --no-e2e --no-variants --steps 20 > gpurun_out/${tag}_ab_fallback.json 2>/dev/null
    a=$(ms gpurun_out/${tag}_ab_intree.json); b=$(ms gpurun_out/${tag}_ab_fallback.json)
    echo "ms per step: in-tree $a, fallback $b"
    if python -c "import sys; sys.exit(0 if float('$b') < 0.995 * float('$a') else 1)"; then
        export SLR_B200_LIB=$PWD/$fb
        echo "fallback $fb (faster: $b vs $a ms)" > gpurun_out/${tag}_lib.txt
    fi
fi
timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
for v in d["config"].get("variants", []): print(v["name"][:40], v["value"], v["roofline_frac"])
print({k: round(v["ms_per_step"], 3) for k, v in d["config"].get("raw_input", {}).items() if isinstance(v, dict)})
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_fused_flow -s 3 -c 1 -f -o gpurun_out/${tag}_flow \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-variants > gpurun_out/${tag}_ncu_full.log 2>&1
timeout 200 python bench_kernels.py --no-cpu > gpurun_out/${tag}_kernel_table.md 2> gpurun_out/${tag}_kernel_table.err
tail -25 gpurun_out/${tag}_kernel_table.md
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-variants > gpurun_out/${tag}_b_ncu.log 2>&1
ls -la gpurun_out | tail -12
