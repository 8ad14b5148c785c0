mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_mf_decode_strict|k3b_code_match|k0_remap_linear' -s 9 -c 6 -o gpurun_out/prof_kernels_r1k python bench_kernels.py --no-cpu --reps 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --gather --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('with_allgather'))"
