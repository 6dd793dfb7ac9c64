#!/bin/bash
# round-2 GPU call 18: single-pass DConv row kernel (frequency branch): operator tests, HTDemucs parity, A/B timing
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_demucs_gpu.py -q > $O/c18_demucs_tests.txt 2>&1; tail -4 $O/c18_demucs_tests.txt | cut -c1-300
B200SEP_DCONV_ROW=0 timeout 300 python tests/dev/demucs_probe.py 4 2>&1 | head -1
PROFILE=1 timeout 300 python tests/dev/demucs_probe.py 4 > $O/c18_htdemucs_profile_b4.txt 2>&1; head -1 $O/c18_htdemucs_profile_b4.txt; grep -E "dconv_row|dconv_k|tc_attention|tc_f32_kernel" $O/c18_htdemucs_profile_b4.txt | cut -c1-70,150-230
timeout 300 python tests/dev/demucs_probe.py 13 2>&1 | head -1
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
