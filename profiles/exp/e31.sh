# session 4: LIMIT experiments on the fp64 M pass (timing only, wrong results): scalar FP64 work next to the DMMA stream
for lib in donk_b200/csrc/libdonk_b200.so profiles/_build/exp_nocentre.so profiles/_build/exp_noscale.so profiles/_build/exp_mboth.so; do
  echo $lib; PROF_ALT_LIB=$lib python profiles/prof_misc.py tc 10000000 72 16 2>&1 | grep "EM iteration (fp64)\|fp64 E pass"
done
