cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t8.log 2>&1; tail -12 gpurun_out/r02_t8.log
for ch in 8 64; do
python bench.py --steps 8 --warmup 3 --chains-per-gpu $ch --no-cpu-baseline --no-size-sweep --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['gpu_launches']); print({k:round(v,2) for k,v in d['fused_step']['phase_ms_per_step'].items()})
    else: print(l.rstrip()[:300])
"
done
SBD_CHAMB_ERRSUB=0 python bench.py --steps 8 --warmup 3 --chains-per-gpu 8 --no-cpu-baseline --no-size-sweep --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('errsub0', d['value'], d['ms_per_step'], {k:round(v,2) for k,v in d['fused_step']['phase_ms_per_step'].items() if 'chamb' in k})
"
