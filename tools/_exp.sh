cd $GRAFT_REPO_ROOT
python tools/run_cman_demo.py --samples 3000 --warmup 2000 --cpu-iters 0 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('default(auto)', round(d['gpu_steps_per_s'],1), round(1e6/d['gpu_steps_per_s'],1),'us/step')
"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests/test_gpu_matlab_dropin.py tests/test_gpu_mex_exec.py -m gpu -q 2>&1 | tail -2
