cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q > gpurun_out/r02_mg2.log 2>&1; tail -5 gpurun_out/r02_mg2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -3 gpurun_out/r02_bench_n2.err | cut -c1-300
python - <<P
import json
for l in open('gpurun_out/r02_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['chains_per_gpu'], d['fused_step']['ms_per_step_serial_order']); print(d['fused_step']['phase_ms_per_step'])
P
