#!/bin/bash
# 2-GPU validation (gpurun --gpus 2): in-library fan-out tests, bench at N = 2 (both arms).
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/multi_tests.log 2>&1; echo "multi rc=$?"; tail -5 gpurun_out/multi_tests.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print({k: d[k] for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches','per_rank_kernel_s_per_step')}, d['e2e'], d['sharded'])" ; tail -3 gpurun_out/bench_n2.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --no-full-run > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"; tail -c 400 gpurun_out/bench_ref_n2.json
