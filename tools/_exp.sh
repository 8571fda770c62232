cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t2.log 2>&1; tail -5 gpurun_out/r02_t2.log
( time python bench.py ) > gpurun_out/r02_b2.json 2> gpurun_out/r02_b2.err; tail -3 gpurun_out/r02_b2.err; python - <<P
import json
for l in open('gpurun_out/r02_b2.json'):
    if l.startswith('{'):
        d=json.loads(l)
        for k,v in d.items(): print(k, ':', json.dumps(v)[:700])
P
( time python bench.py --impl reference --steps 4 --warmup 1 ) 2>&1 | tail -8 | cut -c1-1500
