mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
export TORTTO_B200_LIB=tuning
for rep in 1 2; do for dbg in 0 32; do
TTB_IGEMM_DBG=$dbg timeout 300 python bench.py --cpu-baseline 0 --extra-bf16 0 > gpurun_out/bench_pf$dbg.json 2> gpurun_out/bench_pf.err
tail -1 gpurun_out/bench_pf$dbg.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('r18 dbg=$dbg', 'ms/step %.4f img/s %.0f conv_ms %.3f' % (d['ms_per_step'], d['value'], r['family_ms_per_step']))"
done; done
for dbg in 0 32; do
TTB_IGEMM_DBG=$dbg timeout 300 python bench.py --cpu-baseline 0 --model standard_resnet50 --steps 5 --warmup 3 > gpurun_out/bench_pf50_$dbg.json 2> gpurun_out/bench_pf.err
tail -1 gpurun_out/bench_pf50_$dbg.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('r50 dbg=$dbg', 'ms/step %.4f img/s %.0f conv_ms %.3f' % (d['ms_per_step'], d['value'], r['family_ms_per_step']))"
done
