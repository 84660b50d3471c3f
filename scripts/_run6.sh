mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -q -x -k "stem or bottleneck or resnet50 or replay" > gpurun_out/pytest_stem.log 2>&1; echo "pytest stem rc=$?"
tail -6 gpurun_out/pytest_stem.log
timeout 300 python bench.py --cpu-baseline 0 --model standard_resnet50 --steps 5 --warmup 3 > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err; echo "r50 rc=$?"
tail -1 gpurun_out/bench_r50.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('r50', 'ms/step %.4f img/s %.0f e2e %.0f conv_ms %.3f frac %.3f hbm_ms %.3f hbmfrac %.2f' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['family_ms_per_step'], r['frac'], r['hbm']['family_ms_per_step'], r['hbm']['frac']))
print('   ', {k:v for k,v in list(d['family_ms_per_step']['by_entry_point'].items())[:12]})"
timeout 300 python scripts/per_layer.py standard_resnet50 > gpurun_out/per_layer_standard_resnet50.txt 2>&1
grep "c   3" gpurun_out/per_layer_standard_resnet50.txt; grep "conv total" gpurun_out/per_layer_standard_resnet50.txt
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
