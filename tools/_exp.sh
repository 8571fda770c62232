cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t3.log 2>&1; tail -8 gpurun_out/r02_t3.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-size-sweep --no-extras > gpurun_out/r02_b3.json 2> gpurun_out/r02_b3.err; tail -3 gpurun_out/r02_b3.err
python - <<P
import json
for l in open('gpurun_out/r02_b3.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']); print(d['roofline']['issue_frac'], d['clocks'])
P
SBD_TRACE_HOST=1 python bench.py --steps 20 --warmup 3 --chains-per-gpu 8 --no-cpu-baseline --no-size-sweep --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e'])
    else: print(l.rstrip()[:200])
"
