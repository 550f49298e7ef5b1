cd /root/repo
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_r01_v11_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_r01_v11_n$N.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['e2e']['value'], d['e2e_device_consumer']['value'])
"
tail -3 gpurun_out/bench_n$N.err
