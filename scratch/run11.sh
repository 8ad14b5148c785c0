timeout 900 python bench_kernels.py --no-cpu 2>gpurun_out/kt.err | grep -E "config|kernel \|"
tail -3 gpurun_out/kt.err | cut -c1-300
