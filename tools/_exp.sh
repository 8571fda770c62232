cd $GRAFT_REPO_ROOT
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_n1.json 2>/dev/null
CMD="python bench.py --steps 3 --warmup 3 --chains-per-gpu 8 --no-cpu-baseline --no-size-sweep --no-extras"
$CMD > gpurun_out/plain.log 2>&1 || { echo PLAIN FAILED; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 70 --csv --log-file gpurun_out/r02_launches.csv $CMD > /dev/null 2>&1; echo launches rc=$?
ncu --set full --clock-control none -s 100 -c 24 -f -o /tmp/r02_step $CMD > gpurun_out/ncu_full.log 2>&1; echo full rc=$?
ncu -i /tmp/r02_step.ncu-rep --page raw --csv > gpurun_out/r02_step_raw.csv 2>/dev/null
