mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_r1p.err | tail -1 > gpurun_out/bench_line_r1p.json
cut -c1-330 gpurun_out/bench_line_r1p.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_line_r1p.json
cut -c1-200 gpurun_out/bench_reference_line_r1p.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r1p.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 900 python bench_kernels.py 2> gpurun_out/kernel_table_r1p.jsonl > gpurun_out/kernel_table_r1p.md
cat gpurun_out/kernel_table_r1p.md
