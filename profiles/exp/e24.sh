# session 4: fit_warp_kernel with the fit's sample rows prefetched into L2
for f in 1 0; do echo "DONK_FIT_NO_PREFETCH=$f"; DONK_FIT_NO_PREFETCH=$f python profiles/prof_ilqr.py fit 4096 5 | grep -o "ms_per_launch.*"; DONK_FIT_NO_PREFETCH=$f python profiles/prof_misc.py config2 4096 2>&1 | grep "fit_lr"; DONK_FIT_NO_PREFETCH=$f python profiles/prof_misc.py config1 2>&1 | grep fit_lr; done
python -m pytest tests/test_fit_gmm_kernels.py -m gpu -x -q 2>&1 | tail -2
