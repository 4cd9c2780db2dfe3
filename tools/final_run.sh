#!/bin/bash
# one GPU box: the whole -m gpu suite, the default bench line, the ncu captures the bench's rooflines cite (tag = $1)
t=${1:-r04e}
S=$(date +%s)
python -m pytest tests -q -m gpu > gpurun_out/${t}_gpu_tests.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - S )) s"; tail -n 3 gpurun_out/${t}_gpu_tests.log
python bench.py > gpurun_out/${t}_bench_n1.json 2> gpurun_out/${t}_bench_n1.err; echo "bench rc=$? $(( $(date +%s) - S )) s"
python tools/ray_order_standin.py 2>&1 | tail -n 1 > gpurun_out/${t}_ray_order.log
GPURT_ORDER_MIN_BVH_BYTES=0 python tools/ray_order_standin.py 2>&1 | tail -n 1 >> gpurun_out/${t}_ray_order.log
cat gpurun_out/${t}_ray_order.log
ncu --set full --clock-control none --import-source on -k regex:'k_trace_closest|k_closest_points' -o gpurun_out/${t}_queries -f python tools/ncu_workload.py > gpurun_out/${t}_workload.log 2>&1
ncu -i gpurun_out/${t}_queries.ncu-rep --page raw --csv > gpurun_out/${t}_queries_ncu_raw.csv 2>/dev/null; wc -l gpurun_out/${t}_queries_ncu_raw.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${t}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-strong --no-cpu-baseline > gpurun_out/${t}_bench_under_ncu.log 2>&1
echo "done $(( $(date +%s) - S )) s"
