mkdir -p gpurun_out
TORTTO_B200_DEFER=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_defer1.log 2>&1; echo "pytest defer1 rc=$?"
tail -3 gpurun_out/pytest_gpu_defer1.log
for d in 0 1; do
  TORTTO_B200_DEFER=$d timeout 300 python bench.py --cpu-baseline 0 > gpurun_out/bench_r18_defer$d.json 2> gpurun_out/bench_r18_defer$d.err; echo "r18 defer$d rc=$?"
  TORTTO_B200_DEFER=$d timeout 400 python bench.py --cpu-baseline 0 --model standard_resnet50 --steps 5 --warmup 3 > gpurun_out/bench_r50_defer$d.json 2> gpurun_out/bench_r50_defer$d.err; echo "r50 defer$d rc=$?"
done
for f in gpurun_out/bench_r18_defer0.json gpurun_out/bench_r18_defer1.json gpurun_out/bench_r50_defer0.json gpurun_out/bench_r50_defer1.json; do tail -1 $f | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$f', 'ms/step %.4f img/s %.0f e2e %.0f conv_ms %.3f hbm_ms %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['family_ms_per_step'], r['hbm']['family_ms_per_step']), 'bf16', d.get('bf16',{}).get('ms_per_step'))"; done
# feed experiment (tuning library; WRONG results by design): how much of a layer's time is the weight / activation feed
( export TORTTO_B200_LIB=tuning
for dbg in 0 4 2 1; do echo "== TTB_IGEMM_DBG=$dbg"; TTB_IGEMM_DBG=$dbg timeout 120 python scripts/bench_conv.py; done ) > gpurun_out/dbg_feed.txt 2>&1
cat gpurun_out/dbg_feed.txt
