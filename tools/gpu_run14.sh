mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/pytest.log
(timeout 300 python tools/gpu_parity_probe.py apoa1 nacl dhfr 2>&1 | tail -12) > gpurun_out/probe.log
(timeout 300 python bench.py --steps 6 --warmup 3 2>&1 | grep '^{') > gpurun_out/r02_bench_dhfr.json
(timeout 300 python bench.py --workload apoa1 --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{') > gpurun_out/r02_bench_apoa1.json
(timeout 200 python tools/gpu_fft_vs_cufft.py 2>&1 | tail -8) > gpurun_out/fft.log
cat gpurun_out/pytest.log gpurun_out/probe.log gpurun_out/fft.log
python -c "
import json
for w in ('dhfr','apoa1'):
    j=json.load(open('gpurun_out/r02_bench_%s.json'%w)); print(w, j['value'], j['config']['us_per_md_step'], j['e2e']['value'], j['phases_us'])"
