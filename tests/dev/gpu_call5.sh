#!/bin/bash
# round-2 GPU call 5: conv3 kernel with 16 epilogue warps; cluster 8; graphs in the HTDemucs path
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_mdx_gpu.py tests/test_mdxc_gpu.py -x -q > $O/c5_tests.txt 2>&1; tail -5 $O/c5_tests.txt
for cs in 1 8; do
  B200SEP_CLUSTER=$cs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --also none > $O/c5_bench_cs$cs.json 2> $O/c5_bench_cs$cs.err; tail -2 $O/c5_bench_cs$cs.err
  python - <<PY
import json
d=json.loads(open('$O/c5_bench_cs$cs.json').read().strip().splitlines()[-1])
print('cluster $cs', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['by_category_ms'], round(d['roofline']['achieved'],1), d['clocks'], d['parity'])
PY
done
timeout 900 python -m pytest tests/test_demucs_gpu.py tests/test_roformer_gpu.py tests/test_vr_gpu.py tests/test_sharded_gpu.py -x -q > $O/c5_tests2.txt 2>&1; tail -5 $O/c5_tests2.txt
timeout 600 python bench.py --workload htdemucs_ft --steps 2 --warmup 3 --no-cpu-baseline > $O/c5_bench_demucs.json 2> $O/c5_bench_demucs.err; tail -3 $O/c5_bench_demucs.err; cut -c1-300 $O/c5_bench_demucs.json
B200SEP_GRAPHS=0 timeout 600 python bench.py --workload htdemucs_ft --steps 2 --warmup 3 --no-cpu-baseline > $O/c5_bench_demucs_nograph.json 2> $O/c5_bench_demucs_nograph.err; cut -c1-300 $O/c5_bench_demucs_nograph.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_conv3 -s 0 -c 1 -o $O/r02b_conv3_s0 python tests/dev/dbg_probe.py > $O/c5_ncu_conv3.log 2>&1; tail -2 $O/c5_ncu_conv3.log
ncu -i $O/r02b_conv3_s0.ncu-rep --page details > $O/r02b_conv3_s0_ncu_full.txt 2>&1
grep -E "Duration|TC is|Issue Slots Busy|Registers Per|L2 Cache Throughput|Warp Cycles Per Issued" $O/r02b_conv3_s0_ncu_full.txt | head
