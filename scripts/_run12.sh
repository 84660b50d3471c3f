mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > gpurun_out/pytest_dp.log 2>&1; echo "pytest dp rc=$?"; tail -3 gpurun_out/pytest_dp.log
timeout 300 python bench.py --cpu-baseline 0 --extra-bf16 0 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-baseline 0 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -2 gpurun_out/bench_n2.err
for n in 1 2; do tail -1 gpurun_out/bench_n$n.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N=$n ms/step %.4f img/s %.0f e2e %.0f' % (d['ms_per_step'], d['value'], d['e2e']['value']), 'dp_parity', (d.get('dp_parity') or {}).get('ok'))
print({k:v for k,v in list(d['family_ms_per_step']['by_entry_point'].items())[:10]})"; done
