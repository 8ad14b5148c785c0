#!/bin/bash
# usage: prof_report.sh <rep> <kernel regex in disasm> <rows per launch> <groups file>
rep=$1; pat=$2; rows=$3; groups=$4
ncu -i $rep --page source --csv --print-source sass 2>/dev/null > /tmp/rep_src.csv
( cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all /root/repo/structure-light-reconstructor_b200/libslr_b200.so >/dev/null 2>&1 && nvdisasm -g -c k_fused.sm_100a.cubin > /tmp/rep_disasm.txt )
python $(dirname $0)/sass_groups.py /tmp/rep_src.csv /tmp/rep_disasm.txt "$pat" $rows $groups
python $(dirname $0)/ncu_summary.py $rep '^gpu__time_duration.sum$' '^smsp__inst_executed.sum$' 'smsp__issue_active.avg.pct' 'sm__warps_active.avg.pct_of_peak_sustained_active' 'smsp__thread_inst_executed_per_inst_executed.ratio' '^dram__bytes_(read|write).sum$' 'launch__registers_per_thread$' 'launch__occupancy_limit_(registers|shared_mem)' 'smsp__average_warps_issue_stalled_(wait|barrier|short_scoreboard|long_scoreboard|branch_resolving|math_pipe_throttle|not_selected|mio_throttle|no_instruction)_per_issue_active'
