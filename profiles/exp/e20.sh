# session 4: register budget of ilqr_warp_kernel<20,6> (128 regs at 8 warps x 2 CTAs; 164 at 4 x 3; 208 at 4 x 2)
for lib in donk_b200/csrc/libdonk_b200.so profiles/_build/lb128x3.so profiles/_build/lb128x2.so; do
  for G in 2 4; do echo "G=$G"; DONK_ILQR_WARPS=$G python profiles/prof_altlib.py $lib 1184 5; done
  echo "auto"; python profiles/prof_altlib.py $lib 1184 5
  echo "E=1"; python profiles/prof_altlib.py $lib 4096 5 1
done
