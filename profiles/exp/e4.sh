python profiles/prof_altlib.py donk_b200/csrc/libdonk_b200.so 1184 5
python profiles/prof_altlib.py profiles/_build/pairsplit.so 1184 5
python profiles/prof_altlib.py donk_b200/csrc/libdonk_b200.so 4096 5 1
python profiles/prof_altlib.py profiles/_build/pairsplit.so 4096 5 1
python profiles/prof_misc.py optimize 4096
