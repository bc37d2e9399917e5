run() { echo "== $*"; env "$@" python tools/gpu_iter.py $MODE $SYS 2>&1 | tail -1; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
MODE=timeph SYS=dhfr
run X=1
MODE=timeph SYS=apoa1
run X=1
python tools/gpu_trace.py dhfr 16 | tail -1
