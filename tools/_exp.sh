cd $GRAFT_REPO_ROOT
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
CMD="python bench.py --steps 3 --warmup 3 --chains-per-gpu 8 --no-cpu-baseline --no-size-sweep --no-extras"
$CMD > gpurun_out/plain.log 2>&1 || { echo PLAIN FAILED; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 90 --csv --log-file gpurun_out/r02_launches.csv $CMD > /dev/null 2>&1; echo launches rc=$?
ncu --set full --clock-control none -s 140 -c 32 -f -o /tmp/r02_step $CMD > gpurun_out/ncu_full.log 2>&1; echo full rc=$?
ncu -i /tmp/r02_step.ncu-rep --page raw --csv > gpurun_out/r02_step_raw.csv 2>/dev/null
for ch in 16 32 64; do
  C2="python bench.py --steps 2 --warmup 3 --chains-per-gpu $ch --no-cpu-baseline --no-size-sweep --no-extras"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_chamb_multi -s 30 -c 9 --csv --log-file gpurun_out/r02_chamb_traffic_$ch.csv $C2 > /dev/null 2>&1; echo ch $ch rc=$?
done
