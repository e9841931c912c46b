# session 4: LIMIT experiments on ilqr_warp_kernel<20,6> (results are WRONG by construction; timing only): what would the
# kernel gain if the mirror stores were gone / if the accumulator pair stores cost their conflict-free wavefront count?
for lib in donk_b200/csrc/libdonk_b200.so profiles/_build/exp_nomirror.so profiles/_build/exp_halfpair.so profiles/_build/exp_both.so; do
  python profiles/prof_altlib.py $lib 1184 5 2>&1 | tail -1
  python profiles/prof_altlib.py $lib 4096 5 1 2>&1 | tail -1
done
