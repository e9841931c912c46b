# session 4: fit_lr with prior, cluster weights from one E pass over the transition view vs in-kernel
python -m pytest tests/test_fit_gmm_kernels.py tests/test_full_size_gpu.py tests/test_reference_api.py -m gpu -x -q 2>&1 | tail -3
for f in 1 0; do echo "DONK_FIT_PRIOR_INKERNEL=$f"; DONK_FIT_PRIOR_INKERNEL=$f python profiles/prof_misc.py config2 4096 2>&1 | grep fit_lr; DONK_FIT_PRIOR_INKERNEL=$f python profiles/prof_misc.py config3prior 4096 16 2>&1 | grep fit_lr; DONK_FIT_PRIOR_INKERNEL=$f python profiles/prof_misc.py config3prior 4096 4 2>&1 | grep "prior (20"; done
python profiles/prof_misc.py config5fit 128 500 2>&1 | grep "fit_lr"
