python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_v13.log 2>&1; tail -3 gpurun_out/gpu_tests_v13.log
python tools/diag_ticks.py 220000000 30 > gpurun_out/diag_v13.log 2>&1; cat gpurun_out/diag_v13.log
LPK_PASS_DEBUG=1 python tools/diag_ticks.py 220000000 8 > gpurun_out/diag_v13_nohandler.log 2>&1; cat gpurun_out/diag_v13_nohandler.log
LPK_PASS_DEBUG=3 python tools/diag_ticks.py 220000000 8 > gpurun_out/diag_v13_nohandler_notrial.log 2>&1; cat gpurun_out/diag_v13_nohandler_notrial.log
