# session 4: register-budget experiment, ncu --set full of the round-2 kernels, launch list of the bench command
bash profiles/exp/e20.sh > gpurun_out/e20.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 python profiles/prof_misc.py tc 2000000 72 16 > gpurun_out/plain_tc.log 2>&1 && {
timeout 400 $NCU -k regex:gmm_tc_estep_kernel -s 2 -c 1 -o gpurun_out/prof_tce_s4 python profiles/prof_misc.py tc 2000000 72 16 > gpurun_out/ncu_tce.log 2>&1
timeout 400 $NCU -k regex:gmm_tc_mstep_kernel -s 2 -c 1 -o gpurun_out/prof_tcm_s4 python profiles/prof_misc.py tc 2000000 72 16 > gpurun_out/ncu_tcm.log 2>&1
}
timeout 200 python profiles/prof_misc.py config5 32 100 > gpurun_out/plain_c5.log 2>&1 && {
timeout 400 $NCU -k regex:ilqr_coop_kernel -s 1 -c 1 -o gpurun_out/prof_coop_s4 python profiles/prof_misc.py config5 32 100 > gpurun_out/ncu_c5.log 2>&1
timeout 400 $NCU -k regex:fit_coop_kernel -s 1 -c 1 -o gpurun_out/prof_fitcoop_s4 python profiles/prof_misc.py config5 32 100 > gpurun_out/ncu_c5f.log 2>&1
}
timeout 200 python profiles/prof_ilqr.py fit 1184 3 > gpurun_out/plain_fit.log 2>&1 && \
timeout 400 $NCU -k regex:fit_warp_kernel -s 1 -c 1 -o gpurun_out/prof_fitwarp_s4 python profiles/prof_ilqr.py fit 1184 3 > gpurun_out/ncu_fitw.log 2>&1
python bench.py > gpurun_out/bench_s4c.json 2> gpurun_out/bench_s4c.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:unnamed|donk" --csv --log-file gpurun_out/launches_s4c_bench.csv python bench.py --no-cpu > gpurun_out/ncu_bench_s4c.log 2>&1
ls -la gpurun_out/*.ncu-rep; cat gpurun_out/e20.log
