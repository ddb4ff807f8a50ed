#!/bin/bash
# round 2, GPU call 5 (N GPUs, default 2): process-per-GPU parity tests over CUDA IPC / NVLink, schedule sweep, bench lines
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c5_topo.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_petsc.py tests/test_gpu_15d.py -x -q --tb=short -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/c5_pytest_n$N.log | tail -10
for k in 128 16; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    scripts/r2/mg_sweep.py --gpus $N --k $k --steps 10 2>gpurun_out/c5_sweep_n${N}_k$k.err | grep "^{" | tee gpurun_out/c5_sweep_n${N}_k$k.jsonl | cut -c1-250
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/c5_bench_n$N.err | tail -1 > gpurun_out/c5_bench_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/c5_bench_n$N.json"))
print("N=$N step", round(d["ms_per_step"], 3), "ms", round(d["value"]), "GF; exposed", d.get("exposed_comm_ms"), "compute-only", d.get("compute_only_ms"))
print("roofline", d["roofline"] and (round(d["roofline"]["kernel_ms"], 3), round(d["roofline"]["frac"], 3)), "launches", d["gpu_launches"])
print("e2e", d["e2e"] and {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k in ("ms_per_step", "value", "blocking_ms_per_step", "pcie_GBps_per_gpu_per_direction", "numa")})
print("verified", d["verified"]); print("k16", d["k16"] and (round(d["k16"]["ms_per_step"], 3), round(d["k16"]["value"]), d["k16"]["verified"]))
PY
tail -3 gpurun_out/c5_bench_n$N.err
