timeout 200 python profiles/prof_misc.py tc 1000000 72 16 > gpurun_out/plain_tc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gmm_tc_mstep_kernel -s 1 -c 1 -o gpurun_out/prof_tcm3 python profiles/prof_misc.py tc 1000000 72 16 > gpurun_out/ncu_tcm.log 2>&1
tail -2 gpurun_out/ncu_tcm.log
