# session 4: M pass with the first-moment sums spread over all threads
python -m pytest tests/test_fit_gmm_kernels.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -2
python profiles/prof_misc.py tc 10000000 72 16 2>&1 | grep "EM iteration (fp64)"
python profiles/prof_misc.py tc 4000000 46 16 2>&1 | grep "EM iteration (fp64)"
python profiles/prof_misc.py tc 4000000 35 4 2>&1 | grep "EM iteration (fp64)"
