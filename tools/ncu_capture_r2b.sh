#!/bin/bash
# Round 2, final state (after q_readout was folded into the GEMMs and the K/V half of bridge 1 moved onto tcgen05):
# launch lists + `ncu --set full` captures of the kernels the bench names, exported to CSV (gpurun_out/ <= 64 MiB).
# Every ncu command runs AFTER the same command has exited 0 without ncu.
set -u
OUT=gpurun_out
mkdir -p $OUT
T="python tools/ncu_target.py"
NCU="ncu --clock-control none"
$T > $OUT/r2b_ncu_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
# (1) every launch of two eager forwards (batch 8 x 128^3, bf16) with its device time
$NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $OUT/r2b_launches_forward.csv $T > $OUT/r2b_ncu_l1.log 2>&1
# (2) the bench command itself: CUDA-graph kernel nodes of the first timed step (after 3 warm-up steps + graph captures)
if [ "${NCU_SKIP_BENCH:-0}" != "1" ]; then
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/r2b_ncu_bench_plain.log 2>&1 && \
$NCU --metrics gpu__time_duration.sum --graph-profiling node -s ${NCU_BENCH_SKIP:-44000} -c 2500 --csv \
     --log-file $OUT/r2b_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/r2b_ncu_l2.log 2>&1
fi
# (3) full captures (second forward: skip the first forward's instances)
cap() {  # name regex skip count
  $NCU --set full --import-source on -k "regex:$2" -s $3 -c $4 -o $OUT/r2b_prof_$1 $T > $OUT/r2b_ncu_$1.log 2>&1
  if [ -f $OUT/r2b_prof_$1.ncu-rep ]; then
    ncu -i $OUT/r2b_prof_$1.ncu-rep --page raw --csv > $OUT/r2b_prof_$1_raw.csv 2>/dev/null
    rm -f $OUT/r2b_prof_$1.ncu-rep
  fi
}
# per forward: kv_stream 24 launches (bottleneck 8, bridge 3: 8, bridge 2: 8), kv_project2 8 (bridge 1), linear_tma 96
cap kvp2 "kv_project2_kernel|kvg_combine_kernel" 16 2   # bridge 1: B=8, N=57408, C=128 (kernel + merge)
cap kv8 "kv_stream_kernel" 40 1              # bridge 2: N=10752, C=256
cap lin "linear_tma_kernel" 160 4            # bridge 2, first layer: QKV (+ softmax(Q)) | P W_b^T + LN | FFN1 + GELU | FFN2 + LN
cap ffn "ffn128_kernel" 8 1
cap attnout "attn_out128w_kernel" 8 1
cap tc3 "conv3d_tc3_kernel" 8 6
cap tc "conv3d_tc2_kernel|conv3d_tc_kernel" 14 6
cap halo "conv3d_halo_kernel" 10 4
ls -la $OUT | grep r2b | tail -30
