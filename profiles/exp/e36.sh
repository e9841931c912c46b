# session 4: E pass at d = 144: column blocks in groups of 3 / 6 / 9 (one CTA per SM, 255 registers)
for lib in profiles/_build/ejg3.so donk_b200/csrc/libdonk_b200.so profiles/_build/ejg9.so; do
  echo $lib; PROF_ALT_LIB=$lib python profiles/prof_misc.py epass 8192000 144 16; PROF_ALT_LIB=$lib python profiles/prof_misc.py epass 4000000 100 16
done
python -m pytest tests/test_fit_gmm_kernels.py tests/test_full_size_gpu.py -m gpu -x -q -k "em_ or config5 or prior_weights" 2>&1 | tail -2
