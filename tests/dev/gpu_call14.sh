#!/bin/bash
# round-2 GPU call 14: ncu of the (one-sweep) attention kernel: first launch of a batch-4 HTDemucs forward = frequency self-attention, 2688 x 2688 tokens, 32 (batch, head) pairs
O=gpurun_out/r02; mkdir -p $O
ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_attention -c 1 -o $O/c14_attention python tests/dev/demucs_probe.py 4 > $O/c14_ncu.log 2>&1; tail -2 $O/c14_ncu.log
ncu -i $O/c14_attention.ncu-rep --page details > $O/c14_attention_details.txt 2>/dev/null
grep -E "Duration|Elapsed Cycles|SM Frequency|Executed Ipc|Issue Slots Busy|Registers Per|Dynamic Shared|Theoretical Occ|Achieved Occ|Tensor|No Eligible|Eligible Warps|Issued Warp|Warp Cycles Per Issued|Stall|L1/TEX Hit|Mem Busy|Mem Pipes" $O/c14_attention_details.txt | head -40
ncu -i $O/c14_attention.ncu-rep --page source --csv > $O/c14_attention_source.csv 2>/dev/null; wc -l $O/c14_attention_source.csv
