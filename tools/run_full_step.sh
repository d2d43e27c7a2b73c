#!/bin/bash
# configs[2] at N GPUs: the three timing phases of tools/full_step.py (each its own launch: static-graph DDP), merged into
# gpurun_out/full_step_n${N}_${VARIANT}.json.   usage: tools/run_full_step.sh N [variant] [batch] [steps] [builder]
N=${1:-1}; V=${2:-shared2x2}; B=${3:-64}; S=${4:-8}; BLD=${5:-device}
mkdir -p gpurun_out
for PH in step backbone nosync; do
  if [ "$PH" = nosync ] && [ "$N" = 1 ]; then continue; fi
  if [ "$N" = 1 ]; then
    python bench.py --workload full_step --variant $V --batch $B --phase $PH --builder $BLD --gpus 1 --steps $S --warmup 3 > gpurun_out/fs_${N}_${V}_${PH}.json 2> gpurun_out/fs_${N}_${V}_${PH}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --workload full_step --variant $V --batch $B --phase $PH --builder $BLD --gpus $N --steps $S --warmup 3 > gpurun_out/fs_${N}_${V}_${PH}.json 2> gpurun_out/fs_${N}_${V}_${PH}.err
  fi
  echo "phase $PH rc=$?"
done
python - <<PY
import json
def load(ph):
    try: return json.load(open("gpurun_out/fs_${N}_${V}_%s.json" % ph))
    except Exception as e: return None
s, b, n = load("step"), load("backbone"), load("nosync")
if s:
    s["backbone_only_ms"] = b["ms_per_step"] if b else None
    s["loss_path_share"] = (1 - b["ms_per_step"] / s["ms_per_step"]) if b else None
    s["no_sync_ms"] = n["ms_per_step"] if n else None
    s["exposed_allreduce_share"] = max(0.0, 1 - n["ms_per_step"] / s["ms_per_step"]) if n else 0.0
    json.dump(s, open("gpurun_out/full_step_n${N}_${V}.json", "w"))
    print({k: s[k] for k in ("n_gpus", "value", "ms_per_step", "backbone_only_ms", "loss_path_share", "no_sync_ms", "exposed_allreduce_share")})
PY
