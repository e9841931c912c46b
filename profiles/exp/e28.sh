# session 4: full gpu test run, smoke, bench (N = 1), launch list of the bench command
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s4p.log 2>&1; tail -2 gpurun_out/pytest_gpu_s4p.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s4p.log 2>&1; tail -2 gpurun_out/smoke_s4p.log
python bench.py > gpurun_out/bench_s4p.json 2> gpurun_out/bench_s4p.err; echo bench rc=$?
python bench.py --impl reference > gpurun_out/bench_s4p_reference.json 2> gpurun_out/bench_s4p_reference.err; echo ref rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(brent|cost_l|dyn_log|extended_costs|fit_|fp64_probe|gmm_|ilqr_|kl_action|kmeans_|optimize_ctl|spd_factor|state_fit|to_transitions)" --csv --log-file gpurun_out/launches_s4p_bench.csv python bench.py --no-cpu > gpurun_out/ncu_bench_s4p.log 2>&1; echo ncu rc=$?
wc -l gpurun_out/launches_s4p_bench.csv
