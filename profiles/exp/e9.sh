timeout 300 python -m pytest tests/test_fit_gmm_kernels.py -m gpu -x -q -k "tc or tf32" 2>&1 | tail -3
timeout 200 python profiles/prof_misc.py tc 2000000 72 16 > gpurun_out/plain_tc.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gmm_tc --csv --log-file gpurun_out/launches_tc.csv python profiles/prof_misc.py tc 2000000 72 16 > gpurun_out/ncu_tcl.log 2>&1
tail -3 gpurun_out/plain_tc.log
