cd /root/repo
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py 2>gpurun_out/dist_check.err | tail -1 > gpurun_out/dist_check_n$N.json; cat gpurun_out/dist_check_n$N.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/config_bench_multi.py 2>gpurun_out/config_multi.err | tail -1 > gpurun_out/config_bench_multi_n$N.json; cat gpurun_out/config_bench_multi_n$N.json
tail -5 gpurun_out/config_multi.err
