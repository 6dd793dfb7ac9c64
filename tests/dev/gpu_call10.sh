#!/bin/bash
# round-2 GPU call 10: fused attention (operator test, HTDemucs / Roformer parity with it, A/B timing), Hybrid Demucs v3 tests after the test fixes, LSTM k-split timing
O=gpurun_out/r02; mkdir -p $O
timeout 300 python -m pytest tests/test_tc_f32_gpu.py -q -k fused_attention > $O/c10_attn_tests.txt 2>&1; tail -15 $O/c10_attn_tests.txt | cut -c1-300
timeout 900 python -m pytest tests/test_hdemucs_gpu.py -q > $O/c10_hdemucs_tests.txt 2>&1; tail -15 $O/c10_hdemucs_tests.txt | cut -c1-300
timeout 900 python -m pytest tests/test_demucs_gpu.py tests/test_roformer_gpu.py -q > $O/c10_demucs_roformer_tests.txt 2>&1; tail -8 $O/c10_demucs_roformer_tests.txt | cut -c1-300
for fa in 0 1; do
  B200SEP_FUSED_ATTN=$fa timeout 300 python tests/dev/demucs_probe.py 4 2>&1 | head -1
  B200SEP_FUSED_ATTN=$fa timeout 300 python tests/dev/roformer_probe.py 2 2>&1 | head -1
done
PROFILE=1 timeout 300 python tests/dev/demucs_probe.py 4 > $O/c10_htdemucs_profile_b4_fused.txt 2>&1; sed -n 4,16p $O/c10_htdemucs_profile_b4_fused.txt | cut -c1-60,150-230
PROFILE=1 timeout 300 python tests/dev/hdemucs_probe.py 2 40 > $O/c10_hdemucs_probe.txt 2>&1; head -1 $O/c10_hdemucs_probe.txt; sed -n 6,12p $O/c10_hdemucs_probe.txt | cut -c1-72,150-240
