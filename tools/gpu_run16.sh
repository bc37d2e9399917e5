mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/pytest.log
bash tools/gpu_profile.sh > gpurun_out/profile.log 2>&1
(timeout 300 python bench.py --steps 6 --warmup 3 2>&1 | grep '^{') > gpurun_out/r02_bench_dhfr.json
(B200MD_CLOSE_NM=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{') > gpurun_out/r02_bench_dhfr_fp32only.json
(timeout 300 python bench.py --workload water1m --steps 2 --warmup 3 --md-steps 100 --no-cpu-baseline 2>&1 | grep '^{') > gpurun_out/r02_bench_water1m.json
tail -12 gpurun_out/pytest.log; tail -25 gpurun_out/profile.log
python -c "
import json
for w in ('dhfr','dhfr_fp32only','water1m'):
    try:
        j=json.load(open('gpurun_out/r02_bench_%s.json'%w)); print(w, j['value'], j['config']['us_per_md_step'], j['e2e']['value'], j['roofline']['traffic'], j['phases_us'])
    except Exception as e: print(w, 'failed', e)"
