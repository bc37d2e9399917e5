run() { echo "== $*"; env "$@" python tools/gpu_iter.py $MODE $SYS 2>&1 | tail -1; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
MODE=timeph SYS=dhfr
run X=1
SYS=apoa1
run X=1
SYS=water1m
run X=1
