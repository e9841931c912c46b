# session 4: division-free E pass loader
python -m pytest tests/test_fit_gmm_kernels.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -2
python profiles/prof_misc.py epass 10000000 72 16
python profiles/prof_misc.py epass 8192000 144 16
python profiles/prof_misc.py epass 20480000 35 4
for f in 1 0; do echo "DONK_FIT_PRIOR_INKERNEL=$f"; DONK_FIT_PRIOR_INKERNEL=$f python profiles/prof_misc.py config2 4096 2>&1 | grep "prior (14"; DONK_FIT_PRIOR_INKERNEL=$f python profiles/prof_misc.py config3prior 4096 16 2>&1 | grep "prior (20"; DONK_FIT_PRIOR_INKERNEL=$f python profiles/prof_misc.py config3prior 4096 4 2>&1 | grep "prior (20"; done
python profiles/prof_misc.py config5fit 128 500 2>&1 | grep "fit_lr"
