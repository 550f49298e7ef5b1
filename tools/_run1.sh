cd /root/repo
python -m pytest tests/test_eigensum_gpu.py tests/test_random_phase_gpu.py tests/test_reference_invariants_gpu.py -x -q -m gpu 2>&1 | tail -3
python tools/step_breakdown.py 2048 16384
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tmp.csv python tools/step_breakdown.py 2048 16384 > /dev/null 2>&1
grep -i "spin5" gpurun_out/launches_tmp.csv | tail -4
python bench.py --steps 50 --warmup 3 --no-cpu-baseline --shard-of 8 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['value'], d['e2e_device_consumer']['value'], d['gpu_launches'])
"
