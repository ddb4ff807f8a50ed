#!/bin/bash
# round 2, GPU call 3 (1 GPU): full GPU suite incl. the rank-thread tests of the fused multi-GPU engine, then the new bench.py
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/c3_pytest.log | tail -25
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/c3_bench.err | tail -1 > gpurun_out/c3_bench.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/c3_bench.json"))
print("step", round(d["ms_per_step"], 3), "ms", round(d["value"]), "GF; level-0", round(d["roofline"]["kernel_ms"], 3), "ms frac", round(d["roofline"]["frac"], 3))
print("e2e", d["e2e"] and {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k in ("ms_per_step", "value", "blocking_ms_per_step", "pcie_GBps_per_gpu_per_direction", "numa")})
print("verified", d["verified"])
print("k16", d["k16"] and (round(d["k16"]["ms_per_step"], 3), round(d["k16"]["value"]), d["k16"]["roofline"] and round(d["k16"]["roofline"]["frac"], 3), d["k16"]["verified"]))
print("cpu", d["cpu_baseline"], "launches", d["gpu_launches"], d["clocks"])
PY
tail -5 gpurun_out/c3_bench.err
