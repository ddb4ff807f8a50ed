#!/bin/bash
# round 2, GPU call 10 (1 GPU): the whole GPU suite, smoke(), the default bench line, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 150 --durations=8 > gpurun_out/c10_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/c10_pytest.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/c10_pytest.log | head -20 | cut -c1-200
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/c10_smoke.log
timeout 420 python bench.py 2>gpurun_out/c10_bench.err | tail -1 > gpurun_out/c10_bench_n1.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/c10_bench_n1.json"))
print("step", round(d["ms_per_step"], 3), "ms", round(d["value"]), "GF; level-0", round(d["roofline"]["kernel_ms"], 3), "ms frac", round(d["roofline"]["frac"], 3), "launches", d["gpu_launches"])
print("e2e", d["e2e"] and {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items() if k in ("ms_per_step", "value", "blocking_ms_per_step", "pcie_GBps_per_gpu_per_direction")})
print("verified", d["verified"].get("ok"), d["verified"].get("max_rel_err"), "k16", d["k16"] and (round(d["k16"]["ms_per_step"], 3), round(d["k16"]["value"]), round(d["k16"]["roofline"]["frac"], 3), d["k16"]["verified"].get("ok")))
print("cpu", d["cpu_baseline"] and (round(d["cpu_baseline"]["value"], 1), d["cpu_baseline"]["cores"]), d["clocks"])
PY
tail -2 gpurun_out/c10_bench.err | cut -c1-200
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c10_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-verify --no-k16 > gpurun_out/c10_bench_under_ncu.log 2>&1
grep -c k_spmm gpurun_out/c10_launches_bench.csv
