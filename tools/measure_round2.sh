#!/bin/bash
# Round-2 measurement bundle for one B200 (run under gpurun): tests, ncu captures of the pass per day class, launch list,
# the 2555-tick Nigeria run, sanitizer logs.  Outputs under gpurun_out/ (summaries are copied into profiles/ by hand).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 > $O/r2_m_pytest.log
# full captures of the pass on ticks 20 (campaign), 21 (vital dynamics), 28 (vital dynamics + RI), 40 .. 42 (plain, plain, VD + RI)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tick_pass --launch-skip 19 --launch-count 3 \
    -o $O/prof_r2_v35_sia_vd -f python bench.py --steps 30 --warmup 3 --cpu-agents 200000 --cpu-ticks 2 > $O/r2_m_ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tick_pass --launch-skip 39 --launch-count 3 \
    -o $O/prof_r2_v35_plain -f python bench.py --steps 45 --warmup 3 --cpu-agents 200000 --cpu-ticks 2 > $O/r2_m_ncu_b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_t|k_x|k_births" -c 400 --csv --log-file $O/r2_launches_v35.csv \
    python bench.py --steps 60 --warmup 3 --cpu-agents 200000 --cpu-ticks 2 > $O/r2_m_list.log 2>&1
timeout 900 python bench.py --steps 2555 --warmup 3 > $O/r2_bench_2555.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file $O/r2_sanitizer_memcheck.log python -m pytest \
    "tests/test_gpu_fused.py::test_fused_equals_components_full_feature_set" -x -q -m gpu > $O/r2_m_memcheck_pytest.log 2>&1
timeout 1500 compute-sanitizer --tool racecheck --log-file $O/r2_sanitizer_racecheck.log python -m pytest \
    "tests/test_gpu_fused.py::test_fused_sia_days_small_nodes" -x -q -m gpu > $O/r2_m_racecheck_pytest.log 2>&1
tail -3 $O/r2_m_pytest.log; tail -c 400 $O/r2_bench_2555.log; tail -3 $O/r2_sanitizer_memcheck.log; tail -3 $O/r2_sanitizer_racecheck.log
