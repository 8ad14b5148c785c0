mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_r1m.err | tail -1 > gpurun_out/bench_line_r1m.json
cut -c1-300 gpurun_out/bench_line_r1m.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_line_r1m.json
cut -c1-200 gpurun_out/bench_reference_line_r1m.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r1m.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
tail -3 gpurun_out/launches_bench_r1m.csv | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_mf -s 3 -c 1 -o gpurun_out/prof_fused_strict_r1m python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 900 python bench_kernels.py 2> gpurun_out/kernel_table_r1m.jsonl > gpurun_out/kernel_table_r1m.md
cat gpurun_out/kernel_table_r1m.md
