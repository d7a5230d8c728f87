#!/bin/bash
# Final single-GPU validation: the whole -m gpu suite, smoke(), DRAM traffic of the headline launch, bench both arms.
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_level_engine.py -x -q -k "kats" > gpurun_out/lv_smoke.log 2>&1 || { echo SMOKE FAILED; tail -30 gpurun_out/lv_smoke.log; exit 1; }
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/gputests_final.log 2>&1; echo "rc=$?" >> gpurun_out/gputests_final.log); tail -6 gpurun_out/gputests_final.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_atom.sum,gpu__time_duration.sum --clock-control none -k regex:level_search --launch-skip 3 -c 1 --csv --log-file gpurun_out/r2_level_think0_dram.csv python scripts/prof_level.py 0 exact 2 > gpurun_out/ncu_dram.log 2>&1; tail -8 gpurun_out/r2_level_think0_dram.csv | cut -c1-400
timeout 400 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 300 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
