#!/bin/bash
# A/B of the two sampler kernels (GF_SAMPLER_LANES=16: two rays per warp; default: four lanes per ray) on one GPU:
# parity tests with the default, then a short bench with each.   gpurun --timeout 900 -- 'bash tools/gpu_sampler_ab.sh r02ae'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_ref_kernels.py tests/test_engine_gpu.py -m gpu -x -q \
    -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -5 $OUT/${TAG}_tests.log
for lanes in 4 16; do
  GF_SAMPLER_LANES=$lanes timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-operator-api \
      > $OUT/${TAG}_bench_l${lanes}.json 2> $OUT/${TAG}_bench_l${lanes}.err
  python - <<PY
import json
try:
    b = json.loads(open("$OUT/${TAG}_bench_l${lanes}.json").read().strip().splitlines()[-1])
    print("lanes $lanes ms/step", round(b["ms_per_step"], 3), "rays/s", round(b["value"]), "e2e ms", b["e2e"]["ms_per_step"],
          [(k["kernel"], k["ms"]) for k in b["kernels"][:6]])
except Exception as e:
    print("bench line unreadable:", e); print(open("$OUT/${TAG}_bench_l${lanes}.err").read()[-2000:])
PY
done
