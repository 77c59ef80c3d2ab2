#!/bin/bash
N=${1:-2}; OUT=gpurun_out/r2p_n$N; mkdir -p $OUT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 20 --warmup 5 --developed-steps 0 --no-parity > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "allocs in window", d["device_allocations_in_timed_window_rank0"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["transfer_only_ms_per_step"])
c = d["config4"]; print("config4", c["value"], c["ms_per_step"], c["ms_per_step_rank0"], c["device_allocations_per_step_rank0"], c["clocks"])
print("complete_case", d["complete_case"]["value"], d["complete_case"]["ms_per_step"])
PY
tail -3 $OUT/bench.err | cut -c1-200
