mkdir -p gpurun_out
python scratch/phase_clocks.py strict 2>&1 | tail -9
echo "--- stagger 6000ns"
SLR_FUSED_STAGGER_NS=6000 python scratch/phase_clocks.py strict 2>&1 | tail -9
