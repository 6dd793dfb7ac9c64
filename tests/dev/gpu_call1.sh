#!/bin/bash
# round-2 GPU call 1b: HTDemucs per-kernel profile, baseline bench + batch sweep
O=gpurun_out/r02; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt
PROFILE=1 timeout 600 python tests/dev/demucs_probe.py 4 > $O/demucs_probe_b4.txt 2>&1
PROFILE=1 timeout 600 python tests/dev/demucs_probe.py 1 > $O/demucs_probe_b1.txt 2>&1
timeout 600 python tests/dev/demucs_probe.py 8 > $O/demucs_probe_b8.txt 2>&1
ONCE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/demucs_launches_b4.csv python tests/dev/demucs_probe.py 4 > $O/demucs_ncu.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_base_n1.json 2> $O/bench_base_n1.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch 8 > $O/bench_base_n1_b8.json 2> $O/bench_base_n1_b8.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch 17 > $O/bench_base_n1_b17.json 2> $O/bench_base_n1_b17.err
tail -n 45 $O/demucs_probe_b4.txt; tail -2 $O/demucs_probe_b8.txt; cat $O/bench_base_n1.json $O/bench_base_n1_b8.json $O/bench_base_n1_b17.json | cut -c1-400
