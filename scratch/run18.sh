mkdir -p gpurun_out
timeout 900 python bench_kernels.py 2> gpurun_out/kernel_table_r1n.jsonl > gpurun_out/kernel_table_r1n.md
cat gpurun_out/kernel_table_r1n.md
