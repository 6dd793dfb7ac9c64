#!/bin/bash
# round-2 GPU call 15: attention kernel with sleeping waits (no issue slots for polling warps) and rolled softmax loops (instruction cache)
O=gpurun_out/r02; mkdir -p $O
timeout 300 python -m pytest tests/test_tc_f32_gpu.py -q -k fused_attention > $O/c15_attn_tests.txt 2>&1; tail -3 $O/c15_attn_tests.txt | cut -c1-300
PROFILE=1 timeout 300 python tests/dev/demucs_probe.py 4 > $O/c15_htdemucs_profile_b4.txt 2>&1; head -1 $O/c15_htdemucs_profile_b4.txt; sed -n 6,9p $O/c15_htdemucs_profile_b4.txt | cut -c1-60,150-230
timeout 300 python tests/dev/demucs_probe.py 8 2>&1 | head -1
timeout 300 python tests/dev/roformer_probe.py 2 2>&1 | head -1
timeout 900 python -m pytest tests/test_demucs_gpu.py tests/test_roformer_gpu.py -q > $O/c15_tests.txt 2>&1; tail -3 $O/c15_tests.txt | cut -c1-300
