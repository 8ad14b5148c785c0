timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
b() { timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"; }
b
python scratch/phase_clocks.py strict 2>&1 | tail -9 | head -6
timeout 300 python bench_kernels.py --no-cpu 2>/dev/null | grep -E "k_fused|k3a|config 4"
