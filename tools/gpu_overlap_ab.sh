#!/bin/bash
# A/B of where / how the sampling of the next batch is issued (env knobs of engine.py / bench.py), one GPU.
#   gpurun --timeout 900 -- 'bash tools/gpu_overlap_ab.sh r02am'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
if [ "${TESTS:-1}" = "1" ]; then
timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_ref_kernels.py tests/test_engine_gpu.py -m gpu -x -q \
    -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
fi
i=0
while IFS= read -r cfg; do
  i=$((i+1))
  env $cfg timeout 300 python bench.py --steps ${STEPS:-60} --warmup 10 --no-cpu-baseline --no-operator-api \
      > $OUT/${TAG}_bench_$i.json 2> $OUT/${TAG}_bench_$i.err
  python - <<PY
import json
try:
    b = json.loads(open("$OUT/${TAG}_bench_$i.json").read().strip().splitlines()[-1])
    print("[$cfg] ms/step", round(b["ms_per_step"], 3), "e2e ms", round(b["e2e"]["ms_per_step"], 3),
          [(k["kernel"], k["ms"]) for k in b["kernels"][:6]])
except Exception as e:
    print("[$cfg] bench line unreadable:", e); print(open("$OUT/${TAG}_bench_$i.err").read()[-1500:])
PY
done <<< "${CONFIGS}"
