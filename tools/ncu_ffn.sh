#!/bin/bash
# ncu --set full capture of the fused FFN kernel (tools/ffn_probe.py, first shape), exported to CSV.
set -u
OUT=gpurun_out
mkdir -p $OUT
NAME=${1:-ffn}
python tools/ffn_probe.py > $OUT/ffn_probe.log 2>&1 || { echo "probe failed"; tail -5 $OUT/ffn_probe.log; exit 1; }
ncu --clock-control none --set full --import-source on -k "regex:ffn128_kernel" -s 4 -c 1 -o $OUT/prof_$NAME python tools/ffn_probe.py > $OUT/ncu_$NAME.log 2>&1
ncu -i $OUT/prof_$NAME.ncu-rep --page raw --csv > $OUT/prof_${NAME}_raw.csv 2>/dev/null
ncu -i $OUT/prof_$NAME.ncu-rep --page details --csv > $OUT/prof_${NAME}_details.csv 2>/dev/null
ncu -i $OUT/prof_$NAME.ncu-rep --page source --csv > $OUT/prof_${NAME}_source.csv 2>/dev/null
cat $OUT/ffn_probe.log
