timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench_kernels.py --no-cpu 2> gpurun_out/kernel_table_r1k.jsonl | tee gpurun_out/kernel_table_r1k.md
