for g in 1 5 10; do echo "G=$g"; DONK_ILQR_WARPS=$g python profiles/prof_misc.py optimize 4096; done
echo "auto"; python profiles/prof_misc.py optimize 4096
python profiles/prof_misc.py optimize 3996
python profiles/prof_ilqr.py ilqr 1184 5
python -m pytest tests/test_ilqr_kernels.py -m gpu -x -q | tail -2
