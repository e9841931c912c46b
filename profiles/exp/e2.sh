for g in 2 4 8; do echo "G=$g"; DONK_ILQR_WARPS=$g python profiles/prof_ilqr.py ilqr 1184 5; done
for sk in 4000 9000 14000 18000; do echo "skew=$sk"; DONK_ILQR_SKEW=$sk python profiles/prof_ilqr.py ilqr 1184 5; done
