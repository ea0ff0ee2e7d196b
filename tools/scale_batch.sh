#!/bin/bash
# Scaling runs inside ONE multi-GPU box call:  bash tools/scale_batch.sh "<workload:N:exchange> ..."
#   e.g.  bash tools/scale_batch.sh "cfg2:8:p2p cfg2:8:nccl cfg2:4:p2p cfg4:8:p2p"
set -u
out=gpurun_out/scale; mkdir -p "$out"
port=29520
for spec in $1; do
  IFS=: read -r wl n ex <<< "$spec"
  port=$((port + 1))
  name="${wl}_n${n}_${ex}"
  t0=$SECONDS
  if [ "$n" = "1" ]; then
    TFRS_SHARD_EXCHANGE=$ex timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --workload "$wl" --no-secondary --no-cpu-baseline > "$out/$name.log" 2>&1
  else
    TFRS_SHARD_EXCHANGE=$ex timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus "$n" --steps 20 --warmup 5 --workload "$wl" > "$out/$name.log" 2>&1
  fi
  echo "[$name] rc=$? $((SECONDS - t0))s $(grep -o '"value": [0-9.]*' "$out/$name.log" | head -1) $(grep -o '"ms_per_step": [0-9.]*' "$out/$name.log" | head -1) $(grep -o '"outputs_match_oracle": [a-z]*' "$out/$name.log")"
done
