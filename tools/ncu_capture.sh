#!/bin/bash
# Runs on the GPU box (under gpurun): per-launch time lists and a few `ncu --set full` captures of the
# dominant kernels, exported to CSV so that only small files travel back (gpurun_out/ <= 64 MiB).
set -u
OUT=gpurun_out
mkdir -p $OUT
T="python tools/ncu_target.py"
NCU="ncu --clock-control none"
$T > $OUT/ncu_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
# (1) every launch of two eager forwards (batch 8 x 128^3, bf16) with its device time
$NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $OUT/launches_forward.csv $T > $OUT/ncu_l1.log 2>&1
# (2) full captures, a few launches each (second forward: skip the first forward's instances)
cap() {  # name regex skip count
  $T > /dev/null 2>&1 && $NCU --set full --import-source on -k "regex:$2" -s $3 -c $4 -o $OUT/prof_$1 $T > $OUT/ncu_$1.log 2>&1
  if [ -f $OUT/prof_$1.ncu-rep ]; then
    ncu -i $OUT/prof_$1.ncu-rep --page raw --csv > $OUT/prof_$1_raw.csv 2>/dev/null
    ncu -i $OUT/prof_$1.ncu-rep --page details --csv > $OUT/prof_$1_details.csv 2>/dev/null
    ncu -i $OUT/prof_$1.ncu-rep --page source --csv > $OUT/prof_$1_source.csv 2>/dev/null
    rm -f $OUT/prof_$1.ncu-rep $OUT/prof_$1_details.csv
  fi
}
# per forward the attention kernels run 32 times: bottleneck (8), bridge 3 (8), bridge 2 (8), bridge 1 (8)
cap kv "kv_reduce_mma_kernel" 56 2           # second forward, bridge 1: N=57408 tokens, C=128
cap q "q_readout_mma_kernel" 40 2            # bridge 2 (bridge 1's readout lives in attn_out128_kernel): N=10752, C=256
cap kv8 "kv_reduce_mma_kernel" 48 1          # bridge 2: N=10752, C=256
cap ffn "ffn128_kernel" 8 1                  # fused FFN half, bridge 1 (8 launches per forward)
cap attnout "attn_out128_kernel" 8 1         # fused query half, bridge 1
if [ "${NCU_SKIP_CONV:-0}" != "1" ]; then
cap tc "conv3d_tc2_kernel|conv3d_tc_kernel" 16 6   # six im2col tensor-core conv launches of the second forward
cap tc3 "conv3d_tc3_kernel" 8 6              # six TMA-halo tensor-core conv launches of the second forward
cap halo "conv3d_halo_kernel" 9 3
fi
ls -la $OUT | tail -30
