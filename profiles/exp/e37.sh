python -m pytest tests/test_fit_gmm_kernels.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -2
python profiles/prof_misc.py epass 10000000 72 16
python profiles/prof_misc.py epass 8192000 144 16
python profiles/prof_misc.py epass 20480000 35 4
