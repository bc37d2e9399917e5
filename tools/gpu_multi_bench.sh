# Multi-GPU evidence run (under gpurun --gpus N):  bash tools/gpu_multi_bench.sh N   -> kernel timelines per rank, tests/test_gpu_multi.py, one bench line
N=$1
mkdir -p gpurun_out
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 tools/gpu_trace_multi.py apoa1 8 2>&1 | tail -3) > gpurun_out/trace_x$N.log
(timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -60) > gpurun_out/multi_x$N.log
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus $N --steps 2 --warmup 3 --md-steps 400 --no-cpu-baseline 2>&1 | grep '^{') > gpurun_out/bench_x$N.json
python -c "import sys,json; j=json.load(open('gpurun_out/bench_x$N.json')); print(j['value'], j['config']['us_per_md_step'], 'single', j['single_gpu_same_workload']['us_per_md_step'], 'speedup', j['speedup_vs_single_gpu_same_workload'], j['phases_us'])"
cat gpurun_out/trace_x$N.log gpurun_out/multi_x$N.log
