#!/bin/bash
# round 2, GPU call 6 (N GPUs): fused step with send tiles + copy-engine backward exchange: quick parity, sweep, bench line
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q --tb=short -p no:cacheprovider -k "fused" 2>&1 | tail -4 | tee gpurun_out/c6_pytest_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    scripts/r2/mg_sweep.py --gpus $N --k 128 --steps 10 2>gpurun_out/c6_sweep_n${N}_k128.err | grep "^{" | tee gpurun_out/c6_sweep_n${N}_k128.jsonl | cut -c1-330
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/c6_bench_n$N.err | tail -1 > gpurun_out/c6_bench_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/c6_bench_n$N.json"))
print("N=$N step", round(d["ms_per_step"], 3), "ms", round(d["value"]), "GF; exposed", d.get("exposed_comm_ms"), "compute-only", d.get("compute_only_ms"))
print("roofline", d["roofline"] and (round(d["roofline"]["kernel_ms"], 3), round(d["roofline"]["frac"], 3)), "launches", d["gpu_launches"])
print("e2e", d["e2e"] and {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k in ("ms_per_step", "value", "blocking_ms_per_step", "pcie_GBps_per_gpu_per_direction", "numa")})
print("verified", d["verified"]); print("k16", d["k16"] and (round(d["k16"]["ms_per_step"], 3), round(d["k16"]["value"]), d["k16"]["verified"]))
PY
tail -3 gpurun_out/c6_bench_n$N.err
