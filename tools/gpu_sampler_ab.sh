#!/bin/bash
# A/B of the sampler kernels on one GPU (results are bit-identical): parity tests with the default, then a short bench
# per variant.  VARIANTS = "lanes:groups ..." (GF_SAMPLER_LANES: 16 two rays per warp, -4 quads fused, 4 quads with a
# DFS producer warp; GF_SAMPLER_GROUPS: ray groups per CTA).
#   gpurun --timeout 900 -- 'VARIANTS="4:7 -4:7" bash tools/gpu_sampler_ab.sh r02ak'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_ref_kernels.py tests/test_engine_gpu.py -m gpu -x -q \
    -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -5 $OUT/${TAG}_tests.log
for v in ${VARIANTS:-4:7 -4:7}; do
  lanes=${v%%:*}; groups=${v##*:}
  GF_SAMPLER_LANES=$lanes GF_SAMPLER_GROUPS=$groups timeout 300 python bench.py --steps ${STEPS:-100} --warmup 10 \
      --no-cpu-baseline --no-operator-api > $OUT/${TAG}_bench_l${lanes}_g${groups}.json 2> $OUT/${TAG}_bench_l${lanes}_g${groups}.err
  python - <<PY
import json
try:
    b = json.loads(open("$OUT/${TAG}_bench_l${lanes}_g${groups}.json").read().strip().splitlines()[-1])
    print("lanes $lanes groups $groups ms/step", round(b["ms_per_step"], 3), "rays/s", round(b["value"]), "e2e ms", round(b["e2e"]["ms_per_step"], 3),
          [(k["kernel"], k["ms"]) for k in b["kernels"][:6]])
except Exception as e:
    print("bench line unreadable:", e); print(open("$OUT/${TAG}_bench_l${lanes}_g${groups}.err").read()[-2000:])
PY
done
