python -m pytest tests/test_fit_gmm_kernels.py tests/test_full_size_gpu.py -m gpu -x -q -k "em_ or config4 or gmm" 2>&1 | tail -3
for o in 1 0; do echo "DONK_GMM_M18_OFF=$o"; DONK_GMM_M18_OFF=$o python profiles/prof_misc.py config5fit 32 500 2>&1 | grep "GMM fit"; done
