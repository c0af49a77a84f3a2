#!/bin/bash
OUT=gpurun_out
ncu --clock-control none --set full --import-source on -k "regex:ffn256_kernel" -s 4 -c 1 -o $OUT/prof_ffn256 python tools/ffn_probe.py > $OUT/ncu_ffn256.log 2>&1
ncu -i $OUT/prof_ffn256.ncu-rep --page raw --csv > $OUT/prof_ffn256_raw.csv 2>/dev/null
ncu -i $OUT/prof_ffn256.ncu-rep --page source --csv > $OUT/prof_ffn256_source.csv 2>/dev/null
