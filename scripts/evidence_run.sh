# final evidence of the round on one B200: bench lines of every configuration, launch list, ncu tables, sweep, sanitizer
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_r18.json 2> gpurun_out/bench_r18.err; echo "r18 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python bench.py --cpu-baseline 0 --model standard_resnet50 --steps 5 --warmup 3 > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err; echo "r50 rc=$?"
timeout 300 python bench.py --cpu-baseline 0 --model small_preact_resnet110 > gpurun_out/bench_r110.json 2> gpurun_out/bench_r110.err; echo "r110 rc=$?"
timeout 300 python bench.py --cpu-baseline 0 --model unet --steps 5 --warmup 3 > gpurun_out/bench_unet.json 2> gpurun_out/bench_unet.err; echo "unet rc=$?"
for f in r18 r50 r110 unet; do tail -1 gpurun_out/bench_$f.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$f', 'ms/step %.4f img/s %.0f e2e %.0f conv_ms %.3f frac %.3f hbm_ms %.3f hbmfrac %.2f' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['family_ms_per_step'], r['frac'], r['hbm']['family_ms_per_step'], r['hbm']['frac']), 'bf16', d.get('bf16',{}).get('ms_per_step'), 'traffic', r.get('traffic'))"; done
for m in preact_resnet18 standard_resnet50 unet small_preact_resnet110; do timeout 300 python scripts/per_layer.py $m > gpurun_out/per_layer_$m.txt 2>&1; done
timeout 200 python scripts/per_layer.py preact_resnet18 256 bf16 > gpurun_out/per_layer_preact_resnet18_bf16.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tf32.csv python scripts/one_step.py 2 > gpurun_out/one_step.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_tf32.csv 2 > gpurun_out/launch_summary_tf32.txt; head -12 gpurun_out/launch_summary_tf32.txt
N=$(python - <<'PY'
import csv, re
rows = list(csv.reader(open('gpurun_out/launches_tf32.csv')))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr, data = rows[h], rows[h + 1:]
ki = hdr.index('Kernel Name')
pat = re.compile(r'igemm|wgrad_halo|wgrad_reduce|sum_splits|repack|pack_taps|pack_bf16|stage_channels|unpad_channels')
print(sum(1 for r in data if pat.search(r[ki])) // 2)
PY
)
echo "conv-family launches per step: $N"
REGEX='regex:igemm|wgrad_halo|wgrad_reduce|sum_splits|repack|pack_taps|pack_bf16|stage_channels|unpad_channels'
timeout 900 ncu --set full --clock-control none -k "$REGEX" -s $N -c $N -f -o /tmp/conv_step python scripts/one_step.py 2 > gpurun_out/ncu_full.log 2>&1
python scripts/ncu_table.py /tmp/conv_step.ncu-rep gpurun_out/r2_conv_step_ncu_full.csv gpurun_out/r2_conv_family_traffic.json "ncu --set full --clock-control none -k $REGEX -s $N -c $N python scripts/one_step.py 2  (conv-family launches of ONE preact_resnet18 training step, batch 256, TF32)" | cut -c1-300
timeout 600 ncu --set full --clock-control none -f -o /tmp/hbm python scripts/hbm_kernels.py > gpurun_out/ncu_hbm.log 2>&1
python scripts/ncu_table.py /tmp/hbm.ncu-rep gpurun_out/r2_hbm_kernels_ncu.csv gpurun_out/r2_hbm_kernels_ncu.json "ncu --set full --clock-control none python scripts/hbm_kernels.py" | cut -c1-200
timeout 200 python scripts/hbm_kernels.py --time > gpurun_out/hbm_kernels_time.txt 2>&1; tail -13 gpurun_out/hbm_kernels_time.txt
timeout 900 python scripts/conv_sweep.py gpurun_out/conv_sweep.md > gpurun_out/conv_sweep.log 2>&1; tail -3 gpurun_out/conv_sweep.log
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -x -p no:cacheprovider -k "(test_stem_conv2d_tap_packed_vs_oracle and tf32 and (f7x7 or f7x5)) or (test_conv2d_tensor_path_vs_oracle and tf32 and (n2_c96 or n4_c64_16x16_k64)) or maxpool" > gpurun_out/sanitizer_${tool}_late.log 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_${tool}_late.log; tail -4 gpurun_out/sanitizer_${tool}_late.log
done
