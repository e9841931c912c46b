python -m pytest tests/test_fit_gmm_kernels.py tests/test_full_size_gpu.py tests/test_multi_rank.py -m gpu -x -q -k "kmeans or config4 or prior_init" 2>&1 | tail -3
python profiles/prof_misc.py kmeans 10000000 72 16 2>&1 | tail -4
python profiles/prof_misc.py kmeans 10000000 72 16 2>&1 | tail -3
