timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null
B200MD_USE_GRAPH=0 B200MD_NO_OVERLAP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --md-steps 40 --no-cpu-baseline > gpurun_out/b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pair --launch-skip 5 -c 1 -o gpurun_out/pair_final -f python tools/gpu_iter.py timeph dhfr > gpurun_out/ncu_pair.log 2>&1; tail -1 gpurun_out/ncu_pair.log
B200MD_WORKLOAD=apoa1 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_final_apoa1.json 2>/dev/null
timeout 100 python tools/gpu_trace.py dhfr 16 | tail -1
