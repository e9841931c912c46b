python -m pytest tests/test_reference_api.py -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu > gpurun_out/bench_s4f.json 2> gpurun_out/bench_s4f.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/bench_s4f.json'));g=d['extra']['gps_iteration'];print(g.get('error'));print(g['ms_per_iteration'],g['ms_per_iteration_samples'],g['parts_ms']);print(g['e2e'])"
