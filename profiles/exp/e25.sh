# session 4: fit_warp_kernel register budget: 3 CTAs x 4 warps at 168 registers (96 B of spills) vs 2 CTAs at 255
for i in 1 2; do
python profiles/prof_ilqr.py fit 4096 5 | grep -o "ms_per_launch.*"
PROF_ALT_LIB=profiles/_build/fitlb2.so python profiles/prof_ilqr.py fit 4096 5 | grep -o "ms_per_launch.*"
done
