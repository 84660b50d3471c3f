mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_epilogue.py -m gpu -q -x -k "dgrad_emits" > gpurun_out/pytest_dgradbn.log 2>&1; echo "pytest dgrad_bn rc=$?"
tail -12 gpurun_out/pytest_dgradbn.log
for v in 1 0; do
TORTTO_B200_DGRAD_BN=$v timeout 300 python bench.py --cpu-baseline 0 > gpurun_out/bench_r18_dgradbn$v.json 2> gpurun_out/bench_r18_dgradbn$v.err; echo "r18 rc=$?"
tail -1 gpurun_out/bench_r18_dgradbn$v.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('r18 dgrad_bn=$v', 'ms/step %.4f img/s %.0f e2e %.0f conv_ms %.3f frac %.3f hbm_ms %.3f hbmfrac %.2f' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['family_ms_per_step'], r['frac'], r['hbm']['family_ms_per_step'], r['hbm']['frac']), 'bf16', d.get('bf16',{}).get('ms_per_step'))
print('   ', {k:v for k,v in list(d['family_ms_per_step']['by_entry_point'].items())[:12]})"
done
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --cpu-baseline 0 --model unet --steps 5 --warmup 3 > gpurun_out/bench_unet.json 2> gpurun_out/bench_unet.err; echo "unet rc=$?"
tail -1 gpurun_out/bench_unet.json | cut -c1-200
