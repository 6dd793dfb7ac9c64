#!/bin/bash
# round-2 GPU call 9: Hybrid Demucs v3 parity tests + full-size timing, ncu of the STFT / iSTFT kernels of the MDX path
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_hdemucs_gpu.py -q -x --deselect tests/test_hdemucs_gpu.py::test_full_size_forward_vs_oracle > $O/c9_hdemucs_tests.txt 2>&1; tail -40 $O/c9_hdemucs_tests.txt
timeout 600 python -m pytest tests/test_hdemucs_gpu.py -q -k "full_size or taps" > $O/c9_hdemucs_full.txt 2>&1; tail -25 $O/c9_hdemucs_full.txt
PROFILE=1 timeout 600 python tests/dev/hdemucs_probe.py 2 40 > $O/c9_hdemucs_probe.txt 2>&1; head -40 $O/c9_hdemucs_probe.txt | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stft_forward|istft_frames|istft_ola|demix_ola" -c 4 -o $O/c9_stft python bench.py --minutes 0.5 --steps 1 --warmup 0 --no-cpu-baseline --no-parity --also none > $O/c9_ncu_stft.log 2>&1; tail -3 $O/c9_ncu_stft.log
ncu -i $O/c9_stft.ncu-rep --page details 2>/dev/null | grep -E "^  [a-z_:A-Z]+.*\(|Duration|DRAM Throughput|Memory Throughput|dram__bytes|Registers Per|Achieved Occupancy|L2 Cache Throughput" | head -60
ncu -i $O/c9_stft.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum 2>/dev/null | cut -c1-300 | head -12
