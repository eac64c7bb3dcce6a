#!/bin/bash
# A/B of the streamed-cache mode (mode 3) of the forward tracker: parity tests, then the bench line with and without it.
tag=${1:-rX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-other-rows > gpurun_out/${tag}_bench_stream.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
HSO_TRACK_NO_STREAM=1 timeout 600 python bench.py --no-cpu-baseline --no-other-rows > gpurun_out/${tag}_bench_nostream.json 2>> gpurun_out/${tag}_bench.err; echo "bench exit $?"
python - <<PY
import json
for k in ("stream","nostream"):
    d=json.loads(open(f"gpurun_out/${tag}_bench_{k}.json").read().strip().splitlines()[-1])
    print(k, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["roofline"]["all_levels"]["kernel_ms"], d.get("parity",{}).get("ok"))
PY
