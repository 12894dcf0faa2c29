#!/bin/bash
mkdir -p gpurun_out
for m in fused copy; do
(cd gpurun_out && ISING_B200_HALO=$m timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file launches_d2_$m.csv ../ising_gpu_b200/cuIsing -x 65536 -y 65536 -n 12 -d 2 -t 1.5 > /dev/null 2>&1)
python - gpurun_out/launches_d2_$m.csv $m <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; k=hdr.index('Kernel Name'); v=hdr.index('Metric Value'); d=hdr.index('Device') if 'Device' in hdr else None
agg=collections.defaultdict(list)
for r in rows[1:]: agg[r[k][:60]].append(float(r[v].replace(',','')))
for name,vals in agg.items(): print(sys.argv[2], name, len(vals), 'mean %.1f us'%(sum(vals)/len(vals)/1e3), 'min %.1f max %.1f'%(min(vals)/1e3,max(vals)/1e3))
PY
done
for m in fused copy fused copy; do (cd gpurun_out && ISING_B200_HALO=$m timeout 300 ../ising_gpu_b200/cuIsing -x 65536 -y 65536 -n 128 -d 2 -t 1.5 | tail -1); done
