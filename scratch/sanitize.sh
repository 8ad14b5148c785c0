#!/bin/bash
# compute-sanitizer over the round-2 kernels -> gpurun_out/r2_compute_sanitizer.txt
out=gpurun_out/r2_compute_sanitizer.txt
echo "# compute-sanitizer runs, end of round 2 (B200, sm_100a)" > $out
run() { echo; echo "\$ $*"; timeout 900 "$@" 2>&1 | grep -E "COMPUTE-SANITIZER|passed|failed|error|ERROR SUMMARY|RACECHECK SUMMARY|hazard|smoke ok|Invalid|Race" | head -20; }
{
run compute-sanitizer --tool memcheck python -c 'import __graft_entry__ as g; g.smoke()'
run compute-sanitizer --tool racecheck python -c 'import __graft_entry__ as g; g.smoke()'
run compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_routes.py -m gpu -q -k 'raw or png or merge or band or unaligned or 2048'
run compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_routes.py -m gpu -q -k 'raw or png or merge'
run compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k 'not (full_frame or fuzz)'
run compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k 'fused or rigid or dense or zero_disp or ragged'
run compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_routes.py -m gpu -q -k 'raw or png'
} >> $out
tail -60 $out
