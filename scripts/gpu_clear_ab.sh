# A/B of the clearing-ahead modes (FSGPU_CLEAR_AHEAD): step time = wall clock per operator call, back to back
export PATH=/usr/local/cuda/bin:$PATH
run() { echo "== $*"; env "$@" python scripts/time_ops.py q4rs t3ff 2>&1 | grep "step" | cut -c1-90; }
run FSGPU_CLEAR_AHEAD=0
run FSGPU_CLEAR_AHEAD=2 FSGPU_CLEAR_CTAS=1
run FSGPU_CLEAR_AHEAD=2 FSGPU_CLEAR_CTAS=1 FSGPU_CLEAR_PRIO=1
run FSGPU_CLEAR_AHEAD=2 FSGPU_CLEAR_CTAS=2 FSGPU_CLEAR_PRIO=1
run FSGPU_CLEAR_AHEAD=2 FSGPU_CLEAR_CTAS=4 FSGPU_CLEAR_PRIO=1
