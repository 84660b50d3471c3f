# scaling lines on one 8-GPU box (driver's launch form)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540+n)) bench.py --gpus $n --steps 20 --warmup 5 --cpu-baseline 0 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n$n rc=$?"
tail -1 gpurun_out/bench_n$n.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('N=$n ms/step %.4f img/s %.0f e2e %.0f' % (d['ms_per_step'], d['value'], d['e2e']['value']), 'dp_parity ok:', (d.get('dp_parity') or {}).get('ok'))"
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 --model standard_resnet50 --cpu-baseline 0 > gpurun_out/bench_r50_n8.json 2> gpurun_out/bench_r50_n8.err; echo "r50 n8 rc=$?"
tail -1 gpurun_out/bench_r50_n8.json | cut -c1-260
