#!/bin/bash
# round-2 GPU call 16: attention with pre-split Q / K / V tile images (bulk-copy loader, no producer warps)
O=gpurun_out/r02; mkdir -p $O
timeout 300 python -m pytest tests/test_tc_f32_gpu.py -q -k fused_attention > $O/c16_attn_tests.txt 2>&1; tail -4 $O/c16_attn_tests.txt | cut -c1-300
PROFILE=1 timeout 300 python tests/dev/demucs_probe.py 4 > $O/c16_htdemucs_profile_b4.txt 2>&1; head -1 $O/c16_htdemucs_profile_b4.txt; grep -E "tc_attention|att_pack|tc_f32_kernel" $O/c16_htdemucs_profile_b4.txt | cut -c1-60,150-230
timeout 300 python tests/dev/demucs_probe.py 8 2>&1 | head -1
timeout 300 python tests/dev/roformer_probe.py 2 2>&1 | head -1
timeout 900 python -m pytest tests/test_demucs_gpu.py tests/test_roformer_gpu.py -q > $O/c16_tests.txt 2>&1; tail -3 $O/c16_tests.txt | cut -c1-300
timeout 600 python bench.py --workload htdemucs_ft --no-cpu-baseline --demucs-batch 13 2> $O/c16_bench_b13.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('htdemucs_ft batch 13:', round(d['value'],1), round(d['e2e']['value'],1))"
timeout 600 python bench.py --workload htdemucs_ft --no-cpu-baseline 2> $O/c16_bench_b8.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('htdemucs_ft batch 8:', round(d['value'],1), round(d['e2e']['value'],1))"
