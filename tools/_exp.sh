cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -3 gpurun_out/r02_bench_n8.err
python - <<P
import json
for l in open('gpurun_out/r02_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e'], d['config']['chains_per_gpu'], d['clocks']); print(d.get('config2_laplace_one_image_per_gpu')); print(d['fused_step']['phase_ms_per_step'])
P
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 ) 2>&1 | grep -E "^\{|real" | cut -c1-400
