N=$1; shift
mkdir -p gpurun_out
for k in "$@"; do
  echo "== x$N apoa1 PME_SMS=$k" >> gpurun_out/x$N.log
  (B200MD_PME_SMS=$k timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$((k%10)) bench.py --gpus $N --steps 2 --warmup 3 --md-steps 400 --no-cpu-baseline 2>&1 | grep '^{' | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['config']['us_per_md_step'], 'speedup', j.get('speedup_vs_single_gpu_same_workload'), j['phases_us'])") >> gpurun_out/x$N.log 2>&1
done
cat gpurun_out/x$N.log
