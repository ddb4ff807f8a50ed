#!/bin/bash
# Everything round 1 left gloo-validated only, in one multi-GPU call (N = number of GPUs of the box, >= 2; 4 preferred):
#   gpurun --gpus 4 --timeout 1500 -- 'bash scripts/r2_pending_hardware.sh 4'
# 1. all opt-in hardware tests (banded N-GPU, 1D / 1.5D baselines, split overlap schedule, the two newer golden runs)
# 2. bench.py at N GPUs: overlap 1 vs 2 (device-resident part only)
# 3. arrow vs 1D vs 1.5D on one Barabasi-Albert graph
N=${1:-4}
mkdir -p gpurun_out
export ARROW_TEST_BANDED_GPU=1 ARROW_TEST_PETSC_MULTI_GPU=1 ARROW_TEST_15D_GPU=1 ARROW_TEST_SPLIT_OVERLAP_GPU=1 ARROW_TEST_ALL_GOLDEN_GPU=1
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 | tee gpurun_out/r2_pytest_gpu_all.log | tail -15
for ov in 1 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 5 --overlap $ov --no-e2e --no-cpu 2>gpurun_out/r2_bench_n${N}_ov${ov}.err | tail -1 \
      | tee gpurun_out/r2_bench_n${N}_ov${ov}.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('overlap', $ov, 'ms/step', round(d['ms_per_step'], 3), 'GF', round(d['value']))"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    scripts/compare_baselines.py --vertices 1000000 --neighbors 8 --width 10000 -k 128 --steps 10 --warmup 3 \
    2>gpurun_out/r2_compare.err | tee gpurun_out/r2_compare_n${N}.jsonl
# 4. the k = 16 line of the 10M-row workload (BASELINE.json names k = 16 and 128)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --k 16 --steps 20 --warmup 5 --no-e2e --no-cpu 2>gpurun_out/r2_bench_n${N}_k16.err | tail -1 \
    | tee gpurun_out/r2_bench_n${N}_k16.json | cut -c1-300
timeout 600 python bench.py --k 16 --steps 20 --warmup 5 --no-cpu 2>gpurun_out/r2_bench_n1_k16.err | tail -1 \
    | tee gpurun_out/r2_bench_n1_k16.json | cut -c1-300
