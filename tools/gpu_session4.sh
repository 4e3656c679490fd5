#!/bin/bash
# 1-GPU: validate attention kernels first (short timeouts), then new tests, bench, ncu of the top kernels
mkdir -p gpurun_out
export EPL_ATTENTION=epl
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== attention tests"; timeout -s KILL 240 python -m pytest tests/test_kernels_gpu.py -q -x -k "flash_attention" --timeout 60 2>&1 | tail -15 | tee gpurun_out/pytest_attn.log
echo "== other new tests"; timeout -s KILL 400 python -m pytest tests/test_kernels_gpu.py -q -k "fork or layernorm or linear_and_mlp or gpt2" --timeout 120 2>&1 | tail -8 | tee gpurun_out/pytest_new.log
echo "== bench 1gpu"; timeout -s KILL 400 python bench.py --steps 6 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_v2.log
echo "== bench 1gpu b16"; timeout -s KILL 400 python bench.py --steps 4 --warmup 3 --batch 16 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_b16.log
echo "== launches"; timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
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
