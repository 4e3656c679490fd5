#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== pytest gpu"; timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== bench xl b8"; timeout -s KILL 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench1.log
echo "== launches"; timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_bench.log 2>&1; tail -3 gpurun_out/ncu_bench.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg[r[ki][:70]]+=v; cnt[r[ki][:70]]+=1
tot=sum(agg.values())
with open('gpurun_out/launch_summary.txt','w') as f:
    for k,v in agg.most_common(25):
        line="%6.2f%% %9.1f us %5d  %s"%(100*v/tot,v/1e3,cnt[k],k); print(line); f.write(line+"\n")
PY
