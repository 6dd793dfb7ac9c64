#!/bin/bash
# round-2 GPU call 11: leaner softmax loop of the fused attention (exp2 / MUFU, full-tile fast path), LSTM launch bounds; parity + timing
O=gpurun_out/r02; mkdir -p $O
timeout 300 python -m pytest tests/test_tc_f32_gpu.py -q -k fused_attention > $O/c11_attn_tests.txt 2>&1; tail -5 $O/c11_attn_tests.txt | cut -c1-300
timeout 900 python -m pytest tests/test_hdemucs_gpu.py -q > $O/c11_hdemucs_tests.txt 2>&1; tail -8 $O/c11_hdemucs_tests.txt | cut -c1-300
timeout 900 python -m pytest tests/test_demucs_gpu.py tests/test_roformer_gpu.py -q > $O/c11_demucs_roformer_tests.txt 2>&1; tail -5 $O/c11_demucs_roformer_tests.txt | cut -c1-300
PROFILE=1 timeout 300 python tests/dev/demucs_probe.py 4 > $O/c11_htdemucs_profile_b4.txt 2>&1; head -1 $O/c11_htdemucs_profile_b4.txt; sed -n 4,20p $O/c11_htdemucs_profile_b4.txt | cut -c1-60,150-230
timeout 300 python tests/dev/demucs_probe.py 8 2>&1 | head -1
timeout 300 python tests/dev/roformer_probe.py 2 2>&1 | head -1
PROFILE=1 timeout 300 python tests/dev/hdemucs_probe.py 2 40 > $O/c11_hdemucs_probe.txt 2>&1; head -1 $O/c11_hdemucs_probe.txt; sed -n 6,14p $O/c11_hdemucs_probe.txt | cut -c1-72,150-240
timeout 300 python tests/dev/hdemucs_probe.py 4 40 2>&1 | head -1
