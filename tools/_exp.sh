cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 3 --warmup 3 --chains-per-gpu 8 --no-cpu-baseline --no-size-sweep --no-extras"
$CMD > gpurun_out/plain.log 2>&1 || { echo PLAIN FAILED; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 70 --csv --log-file gpurun_out/r02_launches.csv $CMD > /dev/null 2>&1; echo launches rc=$?
ncu --set full --clock-control none -s 100 -c 24 -f -o /tmp/r02_step $CMD > gpurun_out/ncu_full.log 2>&1; echo full rc=$?
ncu -i /tmp/r02_step.ncu-rep --page raw --csv > gpurun_out/r02_step_raw.csv 2>/dev/null
ls -la /tmp/r02_step.ncu-rep gpurun_out/r02_step_raw.csv
for ch in 16 32 64; do
  C2="python bench.py --steps 2 --warmup 3 --chains-per-gpu $ch --no-cpu-baseline --no-size-sweep --no-extras"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_chamb_multi -s 20 -c 8 --csv --log-file gpurun_out/r02_chamb_traffic_$ch.csv $C2 > /dev/null 2>&1; echo ch $ch rc=$?
done
