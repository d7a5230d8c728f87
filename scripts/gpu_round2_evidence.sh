#!/bin/bash
# One gpurun call: bench (both arms), ncu launch list of the bench command, ncu capture of the level kernel on the
# headline instance, compute-sanitizer on the KAT suites.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_level_engine.py -x -q -k "kats" > gpurun_out/lv_smoke.log 2>&1 || { echo SMOKE FAILED; tail -30 gpurun_out/lv_smoke.log; exit 1; }
echo "== bench ours"; timeout 400 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 600 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
echo "== bench reference"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-sharded --no-cpu-baseline --no-full-run > gpurun_out/launches.log 2>&1; tail -2 gpurun_out/launches.log | cut -c1-300
echo "== ncu level kernel, headline instance"; timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section InstructionStats --clock-control none --import-source on -k regex:level_search -c 1 -o gpurun_out/r2_level_think0_exact python scripts/prof_level.py 0 exact 1 > gpurun_out/ncu_think0.log 2>&1; tail -2 gpurun_out/ncu_think0.log | cut -c1-300
echo "== sanitizer"; 
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_level_engine.py -x -q -k "kats or more_than_32 or wide_keys" > gpurun_out/r2_sanitizer_memcheck_level.log 2>&1; echo "memcheck level rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_level.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "test_kats or set_full_kats or wide_keys or bank_totals" > gpurun_out/r2_sanitizer_memcheck_worklist.log 2>&1; echo "memcheck worklist+scans rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_worklist.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_level_engine.py -x -q -k "kats" > gpurun_out/r2_sanitizer_racecheck_level.log 2>&1; echo "racecheck level rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck_level.log
ls -la gpurun_out | tail -20
