mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --cpu-baseline 0 --model standard_resnet50 --steps 5 --warmup 3 > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err; echo "r50 rc=$?"
timeout 300 python bench.py --cpu-baseline 0 --model unet --steps 5 --warmup 3 > gpurun_out/bench_unet.json 2> gpurun_out/bench_unet.err; echo "unet rc=$?"
timeout 300 python bench.py --cpu-baseline 0 > gpurun_out/bench_r18.json 2> gpurun_out/bench_r18.err; echo "r18 rc=$?"
for f in r18 r50 unet; do tail -1 gpurun_out/bench_$f.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$f', 'ms/step %.4f img/s %.0f e2e %.0f conv_ms %.3f frac %.3f hbm_ms %.3f hbmfrac %.2f' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['family_ms_per_step'], r['frac'], r['hbm']['family_ms_per_step'], r['hbm']['frac']), 'bf16', d.get('bf16',{}).get('ms_per_step'))
print('   ', {k:v for k,v in list(d['family_ms_per_step']['by_entry_point'].items())[:14]})"; done
timeout 200 python scripts/hbm_kernels.py --time > gpurun_out/hbm_kernels_time.txt 2>&1; tail -14 gpurun_out/hbm_kernels_time.txt
