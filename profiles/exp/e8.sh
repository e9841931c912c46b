timeout 300 python -m pytest tests/test_fit_gmm_kernels.py -m gpu -x -q -k "tc or tf32" 2>&1 | tail -3
timeout 300 python profiles/prof_misc.py tc 10000000 72 16 2>&1
