cd $GRAFT_REPO_ROOT
L=semi-blind-image-deblurring-problems-with-tv_b200/lib/libsbd.so
cp $L /tmp/cur.so
run() {
python bench.py --steps 8 --warmup 3 --chains-per-gpu $1 --no-cpu-baseline --no-size-sweep --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], {k:round(v,2) for k,v in d['fused_step']['phase_ms_per_step'].items()})
"
}
for rep in 1 2; do
for v in cur split; do
  if [ $v = cur ]; then cp /tmp/cur.so $L; else cp tools/proto/libsbd_split.so $L; fi
  echo "== $v"; run 8; run 64
done; done
cp tools/proto/libsbd_split.so $L
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t10.log 2>&1; tail -5 gpurun_out/r02_t10.log
cp /tmp/cur.so $L
