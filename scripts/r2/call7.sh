#!/bin/bash
# round 2, GPU call 7 (8 GPUs): parity at 4 ranks, schedule sweep at 8, bench lines at 8 / 4 / 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c7_topo.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q --tb=short -p no:cacheprovider -k "fused" 2>&1 | tail -3 | tee gpurun_out/c7_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    scripts/r2/mg_sweep.py --gpus 8 --k 128 --steps 20 2>gpurun_out/c7_sweep_n8_k128.err | grep "^{" | tee gpurun_out/c7_sweep_n8_k128.jsonl | cut -c1-330
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N \
    bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/c7_bench_n$N.err | tail -1 > gpurun_out/c7_bench_n$N.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/c7_bench_n$N.json"))
    print("N=$N step", round(d["ms_per_step"], 3), "ms", round(d["value"]), "GF; exposed", d.get("exposed_comm_ms"), "compute-only", d.get("compute_only_ms"),
          "roof", d["roofline"] and (round(d["roofline"]["kernel_ms"], 3), round(d["roofline"]["frac"], 3)), "launches", d["gpu_launches"])
    print("   e2e", d["e2e"] and {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k in ("ms_per_step", "value", "blocking_ms_per_step", "pcie_GBps_per_gpu_per_direction", "numa")})
    print("   verified", d["verified"].get("ok"), d["verified"].get("max_rel_err"), "k16", d["k16"] and (round(d["k16"]["ms_per_step"], 3), round(d["k16"]["value"]), d["k16"]["verified"].get("ok")))
except Exception as e:
    print("N=$N failed:", e)
PY
tail -2 gpurun_out/c7_bench_n$N.err | cut -c1-300
done
