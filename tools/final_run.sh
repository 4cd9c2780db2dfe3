#!/bin/bash
# one GPU box: the whole -m gpu suite, the default bench line, the ncu captures the bench's rooflines cite, the launch list of
# the bench and the compute-sanitizer passes (tag = $1); every step under its own timeout so that one stuck step cannot eat the box
t=${1:-r05e}
S=$(date +%s)
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/${t}_gpu_tests.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - S )) s"; tail -n 3 gpurun_out/${t}_gpu_tests.log
timeout 200 python bench.py > gpurun_out/${t}_bench_n1.json 2> gpurun_out/${t}_bench_n1.err; echo "bench rc=$? $(( $(date +%s) - S )) s"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_trace_closest|k_closest_points' -o gpurun_out/${t}_queries -f python tools/ncu_workload.py > gpurun_out/${t}_workload.log 2>&1
ncu -i gpurun_out/${t}_queries.ncu-rep --page raw --csv > gpurun_out/${t}_queries_ncu_raw.csv 2>/dev/null; wc -l gpurun_out/${t}_queries_ncu_raw.csv
echo "ncu full $(( $(date +%s) - S )) s"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${t}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-strong --no-cpu-baseline > gpurun_out/${t}_bench_under_ncu.log 2>&1
echo "launch list $(( $(date +%s) - S )) s"
for tool in memcheck racecheck synccheck initcheck; do
  timeout 60 compute-sanitizer --tool $tool python tools/sanitize_workload.py --no-gather > gpurun_out/${t}_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$? $(tail -n 1 gpurun_out/${t}_sanitizer_$tool.txt) $(( $(date +%s) - S )) s"
done
echo "done $(( $(date +%s) - S )) s"
