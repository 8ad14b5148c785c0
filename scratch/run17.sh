python scratch/k0_only.py
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k0_remap -s 3 -c 1 -o gpurun_out/prof_k0_r1n python scratch/k0_only.py > /dev/null 2>&1
ls -la gpurun_out/prof_k0_r1n.ncu-rep
