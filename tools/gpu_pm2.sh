#!/bin/bash
mkdir -p gpurun_out
M=sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed.sum,smsp__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
for b in "$@"; do
timeout 200 tools/$b > gpurun_out/$b.txt 2>&1; cat gpurun_out/$b.txt
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${b}_ncu.csv tools/$b > /dev/null 2>&1
python - gpurun_out/${b}_ncu.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); idi=hdr.index('ID')
d={}
for r in rows[1:]: d.setdefault((int(r[idi]),r[ki]),{})[r[mi]]=float(r[vi].replace(',',''))
for (i,k),m in sorted(d.items()):
    if i%2==0: continue
    tot=m['sm__inst_executed.sum']; cyc=m['smsp__cycles_active.avg']
    print(f"{k[:40]:40s} IPC {tot/592/cyc:.3f} alu% {m['sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed']:.1f} heavy% {m['sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed']:.1f} | instr alu {m['sm__inst_executed_pipe_alu.sum']/tot:.2f} heavy {m['sm__inst_executed_pipe_fmaheavy.sum']/tot:.2f} | disp {m['smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio']:.2f} math {m['smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']:.2f} wait {m['smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']:.2f}")
PY
done
