mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest.log
(timeout 300 python tools/gpu_parity_probe.py apoa1 2>&1 | tail -8) > gpurun_out/probe.log
for w in dhfr apoa1; do for k in 0 24 36 48 64; do
  echo "== $w PME_SMS=$k" >> gpurun_out/sms.log
  (B200MD_PME_SMS=$k timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --via cabi 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['config']['us_per_md_step'], j['e2e']['value'], j['phases_us'])") >> gpurun_out/sms.log 2>&1
done; done
cat gpurun_out/pytest.log gpurun_out/probe.log gpurun_out/sms.log
