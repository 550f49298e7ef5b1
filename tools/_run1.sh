cd /root/repo
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r01_v10.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pairwise_kernel -s 3 -c 1 -o gpurun_out/prof_pairwise_r01_v10 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | tail -5
