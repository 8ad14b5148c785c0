import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
h=rows[0]; u=rows[1]; v=rows[2]
pats = sys.argv[2:] or ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput','sm__throughput.avg.pct','launch__registers','launch__occupancy','warps_active.avg.pct','smsp__inst_executed.sum','issue_active.avg.pct','bank_conflicts','warp_issue_stalled.*per_warp_active','launch__grid_size','launch__block_size','sm__inst_executed_pipe','smsp__thread_inst_executed_per_inst','lts__t_bytes.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__pipe.*cycles_active.avg.pct','smsp__average_warp']
import re
for i,n in enumerate(h):
    if any(re.search(p,n) for p in pats): print(f"{n:90s} {u[i]:14s} {v[i]}")
