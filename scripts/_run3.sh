# 2-GPU validation: DP parity under pytest, then the scaling bench line at N = 2 (driver's launch form)
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > gpurun_out/pytest_dp.log 2>&1; echo "pytest dp rc=$?"; tail -5 gpurun_out/pytest_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/bench_n2.err
tail -1 gpurun_out/bench_n2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('N=2 ms/step %.4f img/s %.0f e2e %.0f conv_ms %.3f hbm_ms %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], r['family_ms_per_step'], r['hbm']['family_ms_per_step']))
print('dp_parity', d.get('dp_parity'))
print({k:v for k,v in list(d['family_ms_per_step']['by_entry_point'].items())[:14]})"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --model standard_resnet50 --cpu-baseline 0 > gpurun_out/bench_r50_n2.json 2> gpurun_out/bench_r50_n2.err; echo "bench r50 n2 rc=$?"
tail -1 gpurun_out/bench_r50_n2.json | cut -c1-400
