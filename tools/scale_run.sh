#!/bin/bash
# 1/2/4/8-GPU runs of the three bench workloads (cfg5 sweep on the 128x128 mesh, cfg3 GD step, cfg4 ensemble) on one box (run under: gpurun --gpus 8 -- bash tools/scale_run.sh)
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N))"; fi
  $L bench.py --gpus $N --workload sweep --steps 5 --warmup 3 2>gpurun_out/sweep_n$N.err | grep '^{' > gpurun_out/sweep_n$N.json
  $L bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-sweep 2>gpurun_out/cfg3_n$N.err | grep '^{' > gpurun_out/cfg3_n$N.json
  $L bench.py --gpus $N --workload ensemble --steps 20 --warmup 3 2>gpurun_out/ensemble_n$N.err | grep '^{' > gpurun_out/ensemble_n$N.json
  python - <<PY
import json
for w in ("sweep","cfg3","ensemble"):
    try:
        d=json.load(open(f"gpurun_out/{w}_n$N.json")); print(w,"N=$N","ms",round(d["ms_per_step"],3),"value",f'{d["value"]:.4g}')
    except Exception as e: print(w,"N=$N","FAILED",e)
PY
done
