#!/bin/bash
# DRAM bytes of every GEMM launch of one step (ncu metrics-only pass) -> gpurun_out/gemm_dram.csv; tools/ncu_traffic.py sums it.
O=gpurun_out; mkdir -p $O
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
  -k regex:gemm_tc_kernel --csv --log-file $O/gemm_dram.csv python tools/ncu_step.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/gemm_dram.csv') if not l.startswith('==')]
tot={}; n=set()
for r in csv.DictReader(rows):
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']; m=r['Metric Name']
    scale={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9,'ns':1e-9,'us':1e-6,'ms':1e-3}.get(u,1)
    tot[m]=tot.get(m,0)+v*scale; n.add(r['ID'])
print('gemm launches', len(n), {k: v for k,v in tot.items()})
print('dram bytes per launch', (tot['dram__bytes_read.sum']+tot['dram__bytes_write.sum'])/len(n))
PY
