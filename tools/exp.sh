run() { echo "== $*"; env "$@" timeout 120 python tools/gpu_iter.py $MODE $SYS 2>&1 | tail -1; }
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
MODE=timeph SYS=dhfr
run X=1
MODE=time SYS=apoa1
run X=1
