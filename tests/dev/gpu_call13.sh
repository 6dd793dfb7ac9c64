#!/bin/bash
# round-2 GPU call 13: two-sweep attention kernel, conv_gen staging; parity + timing
O=gpurun_out/r02; mkdir -p $O
timeout 300 python -m pytest tests/test_tc_f32_gpu.py -q -k fused_attention > $O/c13_attn_tests.txt 2>&1; tail -5 $O/c13_attn_tests.txt | cut -c1-300
timeout 900 python -m pytest tests/test_demucs_gpu.py tests/test_roformer_gpu.py tests/test_hdemucs_gpu.py tests/test_vr_gpu.py -q > $O/c13_tests.txt 2>&1; tail -5 $O/c13_tests.txt | cut -c1-300
PROFILE=1 timeout 300 python tests/dev/demucs_probe.py 4 > $O/c13_htdemucs_profile_b4.txt 2>&1; head -1 $O/c13_htdemucs_profile_b4.txt; sed -n 6,22p $O/c13_htdemucs_profile_b4.txt | cut -c1-60,150-230
timeout 300 python tests/dev/demucs_probe.py 8 2>&1 | head -1
timeout 300 python tests/dev/roformer_probe.py 2 2>&1 | head -1
