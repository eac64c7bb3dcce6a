run() { echo "== $1"; shift; env "$@" timeout 400 python bench.py --no-cpu-baseline --no-other-rows --no-parity-check --no-e2e --steps 6 $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['roofline']['all_levels']['kernel_ms']); print({k:round(v,4) for k,v in d['diag'].items() if 'frac' in k or 'control' in k})"; }
EXTRA="" run "fwd B=2368" X=1
EXTRA="--ic" run "IC B=2368" X=1
