cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -2 gpurun_out/r02_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
python - <<P
import json
for n in (2,4,8):
    for l in open(f'gpurun_out/r02_bench_n{n}.json'):
        if l.startswith('{'):
            d=json.loads(l); print(n, round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['chains_per_gpu'], d['clocks']['sm_mhz'], d.get('config2_laplace_one_image_per_gpu',{}).get('image_steps_per_s_device'))
P
