#!/bin/bash
# usage: scripts/scale_run.sh "<list of N>" [extra bench args]   -> gpurun_out/scale_<N>gpu.json
NS="$1"; shift
for n in $NS; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 5 --no-cpu "$@" > gpurun_out/scale_${n}gpu.json 2> gpurun_out/scale_${n}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 5 "$@" > gpurun_out/scale_${n}gpu.json 2> gpurun_out/scale_${n}gpu.err
  fi
  echo "N=$n rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_${n}gpu.json"))
    print("  ms/step %.2f  value %.3f G/s  e2e %.3f G/s  cycles/step %.1f launches %d clocks %s" % (d["ms_per_step"], d["value"]/1e9, d["e2e"]["value"]/1e9, d["config"]["mg_cycles_per_step"], d["gpu_launches"], d["clocks"]))
    print("  kernels", d["kernel_ms_per_step"], "sum %.2f" % sum(d["kernel_ms_per_step"].values()))
except Exception as e:
    print("  no json:", e)
PY
  tail -2 gpurun_out/scale_${n}gpu.err
done
