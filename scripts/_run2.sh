mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_r18.json 2> gpurun_out/bench_r18.err; echo "r18 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
timeout 300 python bench.py --cpu-baseline 0 --model standard_resnet50 --steps 5 --warmup 3 > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err; echo "r50 rc=$?"
timeout 300 python bench.py --cpu-baseline 0 --model small_preact_resnet110 > gpurun_out/bench_r110.json 2> gpurun_out/bench_r110.err; echo "r110 rc=$?"
timeout 300 python bench.py --cpu-baseline 0 --model unet --steps 5 --warmup 3 > gpurun_out/bench_unet.json 2> gpurun_out/bench_unet.err; echo "unet rc=$?"
for f in r18 r50 r110 unet; do tail -1 gpurun_out/bench_$f.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$f', 'ms/step %.4f img/s %.0f e2e %.0f conv_ms %.3f frac %.3f hbm_ms %.3f hbmfrac %.2f' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['family_ms_per_step'], r['frac'], r['hbm']['family_ms_per_step'], r['hbm']['frac']), 'bf16', d.get('bf16',{}).get('ms_per_step'), 'cpu', d['cpu_baseline'])"; done
# launch list of a step (2 steps profiled, the last one summarised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tf32.csv python scripts/one_step.py 2 > gpurun_out/one_step.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_tf32.csv 2 > gpurun_out/launch_summary_tf32.txt; head -40 gpurun_out/launch_summary_tf32.txt
N=$(python - <<'PY'
import csv, re
rows = list(csv.reader(open('gpurun_out/launches_tf32.csv')))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr, data = rows[h], rows[h + 1:]
ki = hdr.index('Kernel Name')
pat = re.compile(r'igemm|wgrad_halo|wgrad_reduce|sum_splits|repack|pack_taps|pack_bf16|stage_channels|unpad_channels')
n = sum(1 for r in data if pat.search(r[ki]))
print(n // 2)
PY
)
echo "conv-family launches per step: $N"
REGEX='regex:igemm|wgrad_halo|wgrad_reduce|sum_splits|repack|pack_taps|pack_bf16|stage_channels|unpad_channels'
timeout 900 ncu --set full --clock-control none -k "$REGEX" -s $N -c $N -f -o /tmp/conv_step python scripts/one_step.py 2 > gpurun_out/ncu_full.log 2>&1
python scripts/ncu_table.py /tmp/conv_step.ncu-rep gpurun_out/r2_conv_step_ncu_full.csv gpurun_out/r2_conv_family_traffic.json "ncu --set full --clock-control none -k $REGEX -s $N -c $N python scripts/one_step.py 2  (conv-family launches of ONE preact_resnet18 training step, batch 256, TF32)"
