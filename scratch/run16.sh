timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench_kernels.py --no-cpu 2>/dev/null | grep -E "k0_rectify"
