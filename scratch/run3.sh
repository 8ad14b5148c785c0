mkdir -p gpurun_out
export SLR_B200_LIB=$PWD/structure-light-reconstructor_b200/libslr_b200.so
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_mf -s 2 -c 1 -o gpurun_out/prof_fused_corrected_r1i python scratch/phase_clocks.py corrected > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_mf -s 2 -c 1 -o gpurun_out/prof_fused_strict_r1i python scratch/phase_clocks.py strict > /dev/null 2>&1
ls -la gpurun_out | tail -4
