#!/bin/bash
# Final single-GPU validation: the whole -m gpu suite, smoke(), crash-heavy timing (warm context), bench (ours).
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_level_engine.py -x -q -k "kats" > gpurun_out/lv_smoke.log 2>&1 || { echo SMOKE FAILED; tail -30 gpurun_out/lv_smoke.log; exit 1; }
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gputests_final.log 2>&1; echo "rc=$?" >> gpurun_out/gputests_final.log); tail -12 gpurun_out/gputests_final.log | cut -c1-300
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python scripts/crashy_valid.py 2 > gpurun_out/crashy.log 2>&1; grep -v "^   scouts" gpurun_out/crashy.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ours.json')); print({k: d[k] for k in ('value','ms_per_step','time_to_verdict_s','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['verdict_to_verdict']['time_to_verdict_s'])"; tail -3 gpurun_out/bench_ours.err
