#!/bin/bash
# Round-end record on one B200: GPU tests, the bench line (with CPU baselines), the reference arm,
# the single-posterior latency, the ncu launch list of bench.py itself, DRAM traffic per kernel and
# a full capture of the dominant kernel.  Everything lands in gpurun_out/ under r02_*_final names.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r02_pytest_gpu_final.log; cat $O/r02_pytest_gpu_final.log
python bench.py > $O/r02_bench_n1_final.json 2> $O/r02_bench_n1_final.err; tail -2 $O/r02_bench_n1_final.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_final.json 2> $O/r02_bench_reference_final.err
python tools/latency_probe.py > $O/r02_latency_c1_final.json 2>&1
B="python bench.py --steps 2 --warmup 3 --resilience-m 0 --gcp-candidates 0 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_launches_bench_final.csv $B > $O/r02_ncu_bench_final.log 2>&1
python tools/launch_summary.py $O/r02_launches_bench_final.csv > $O/r02_launches_bench_final_summary.txt
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_traffic_final.csv python tools/profile_step.py 10000 2 > $O/r02_ncu_traffic_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rhfq_saq_post -c 1 -s 1 -f -o $O/r02_saq_post_final python tools/profile_step.py 10000 2 > $O/r02_ncu_saq_post_final.log 2>&1
tail -2 $O/r02_ncu_saq_post_final.log
