for m in 0 1 2 4 8 14; do
  LTU_KVP_MODE=$m ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_kvp_m$m.csv python tools/kvp_probe.py > /dev/null 2>&1
  echo "mode $m: $(grep kv_project2 gpurun_out/r3_kvp_m$m.csv | tail -2 | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done
