N=$1; shift
mkdir -p gpurun_out
(B200MD_PME_SMS=$TRACE_SMS timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tools/gpu_trace_multi.py apoa1 8 2>&1 | tail -3) > gpurun_out/trace_x$N.log
for k in "$@"; do
  echo "== x$N apoa1 PME_SMS=$k" >> gpurun_out/x$N.log
  (B200MD_PME_SMS=$k timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 296$N$((k%10)) bench.py --gpus $N --steps 2 --warmup 3 --md-steps 400 --no-cpu-baseline 2>&1 | grep '^{' | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['config']['us_per_md_step'], 'single', j['single_gpu_same_workload']['us_per_md_step'], j['phases_us'])") >> gpurun_out/x$N.log 2>&1
done
cat gpurun_out/trace_x$N.log gpurun_out/x$N.log
