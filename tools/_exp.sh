cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q -k "cufft or fft2 or blur or sapg_sizes or golden or likelihood" > gpurun_out/r02_t12.log 2>&1; tail -4 gpurun_out/r02_t12.log
CMD="python bench.py --steps 6 --warmup 3 --chains-per-gpu 8 --no-cpu-baseline --no-size-sweep --no-extras"
$CMD 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step']); print(d['fused_step']['phase_ms_per_step'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_cols|k_rows' -s 20 -c 30 --csv --log-file gpurun_out/r02_l8.csv $CMD > /dev/null 2>&1
python - <<P
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_l8.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); 
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki][:60]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in d.items(): print(k, len(v), round(sum(v)/len(v)/1e3,1),'us', 'max',round(max(v)/1e3,1))
P
