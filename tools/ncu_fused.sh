#!/bin/bash
# ncu --set full captures of the two fused transformer kernels (tools/ffn_probe.py, tools/attn_probe.py; first shape:
# 8 x 57408 tokens, d_model 128), exported to CSV.
set -u
OUT=gpurun_out
mkdir -p $OUT
python tools/ffn_probe.py > $OUT/ffn_probe.log 2>&1 || { echo "ffn probe failed"; tail -5 $OUT/ffn_probe.log; exit 1; }
python tools/attn_probe.py > $OUT/attn_probe.log 2>&1 || { echo "attn probe failed"; tail -5 $OUT/attn_probe.log; exit 1; }
cap() {  # name regex skip script
  ncu --clock-control none --set full --import-source on -k "regex:$2" -s $3 -c 1 -o $OUT/prof_$1 python $4 > $OUT/ncu_$1.log 2>&1
  ncu -i $OUT/prof_$1.ncu-rep --page raw --csv > $OUT/prof_$1_raw.csv 2>/dev/null
  ncu -i $OUT/prof_$1.ncu-rep --page details --csv > $OUT/prof_$1_details.csv 2>/dev/null
  ncu -i $OUT/prof_$1.ncu-rep --page source --csv > $OUT/prof_$1_source.csv 2>/dev/null
}
cap ffn "ffn128_kernel" 4 tools/ffn_probe.py
cap attnout "attn_out128_kernel" 4 tools/attn_probe.py
cat $OUT/ffn_probe.log $OUT/attn_probe.log
