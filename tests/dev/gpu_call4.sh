#!/bin/bash
# round-2 GPU call 4: GEMM activation multicast sweep, config-1 test, ncu of the conv3 kernel, launch list of the bench
O=gpurun_out/r02; mkdir -p $O
B200SEP_CLUSTER_GEMM=4 B200SEP_CLUSTER=2 timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_mdx_gpu.py -x -q > $O/c4_tests_cluster.txt 2>&1; tail -5 $O/c4_tests_cluster.txt
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_mdx_gpu.py tests/test_mdxc_gpu.py tests/test_roformer_gpu.py -x -q > $O/c4_tests.txt 2>&1; tail -5 $O/c4_tests.txt
for cg in 1 2 4; do
  B200SEP_CLUSTER_GEMM=$cg timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --also none > $O/c4_bench_cg$cg.json 2> $O/c4_bench_cg$cg.err; tail -2 $O/c4_bench_cg$cg.err
  python - <<PY
import json
d=json.loads(open('$O/c4_bench_cg$cg.json').read().strip().splitlines()[-1])
print('gemm cluster $cg', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['by_category_ms'], round(d['roofline']['achieved'],1), d['clocks'], d['parity'])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_conv3 -s 0 -c 1 -o $O/r02_conv3_s0 python tests/dev/dbg_probe.py > $O/c4_ncu_conv3.log 2>&1; tail -3 $O/c4_ncu_conv3.log
ncu -i $O/r02_conv3_s0.ncu-rep --page details > $O/r02_conv3_s0_ncu_full.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_pair -s 3 -c 1 -o $O/r02_tdf_gemm python tests/dev/dbg_probe.py > $O/c4_ncu_gemm.log 2>&1; tail -3 $O/c4_ncu_gemm.log
ncu -i $O/r02_tdf_gemm.ncu-rep --page details > $O/r02_tdf_gemm_ncu_full.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches.csv python bench.py --minutes 0.5 --steps 1 --warmup 1 --no-cpu-baseline --no-parity --also none > $O/c4_launch.log 2>&1; tail -2 $O/c4_launch.log | cut -c1-300
grep -E "Duration|Tensor|Issue Slots Busy|Registers|DRAM Throughput|L2 Cache Throughput|Eligible|No Eligible|Executed Ipc" $O/r02_conv3_s0_ncu_full.txt | head -30
