python -m pytest tests/test_fit_gmm_kernels.py -m gpu -x -q -k "kmeans" 2>&1 | tail -8
for s in 1 0; do echo "DONK_KMEANS_SCALAR=$s"; DONK_KMEANS_SCALAR=$s python profiles/prof_misc.py kmeans 10000000 72 16 2>&1 | tail -4; done
