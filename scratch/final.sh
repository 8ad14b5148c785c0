mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_r1q.err | tail -1 > gpurun_out/bench_line_r1q.json
cut -c1-330 gpurun_out/bench_line_r1q.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_line_r1q.json
cut -c1-200 gpurun_out/bench_reference_line_r1q.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r1q.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_mf -s 3 -c 1 -o gpurun_out/prof_fused_strict_r1q python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 900 python bench_kernels.py 2> gpurun_out/kernel_table_r1q.jsonl > gpurun_out/kernel_table_r1q.md
cat gpurun_out/kernel_table_r1q.md
