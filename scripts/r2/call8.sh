#!/bin/bash
# round 2, GPU call 8 (4 GPUs): rotated push / pull order -- parity through bench's rank-1 property, phases, step time
N=${1:-4}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-k16 2>gpurun_out/c8_bench_n$N.err | tail -1 > gpurun_out/c8_bench_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/c8_bench_n$N.json"))
print("N=$N step", round(d["ms_per_step"], 3), "ms", round(d["value"]), "GF; exposed", d.get("exposed_comm_ms"), "compute-only", d.get("compute_only_ms"), "verified", d["verified"].get("ok"), d["verified"].get("max_rel_err"))
PY
tail -2 gpurun_out/c8_bench_n$N.err | cut -c1-300
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
    scripts/r2/mg_sweep.py --gpus $N --k 128 --steps 20 2>gpurun_out/c8_sweep_n${N}_k128.err | grep "^{" | tee gpurun_out/c8_sweep_n${N}_k128.jsonl | cut -c1-330
