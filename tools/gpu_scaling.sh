#!/bin/bash
# Multi-GPU bench runs of one round on an 8-GPU box: gpurun --gpus 8 -- 'bash tools/gpu_scaling.sh r02m'
# global stage at N = 8, 4, 2 (weak scaling, peer-memory exchange), the same at N = 8 through NCCL (GF_PEER_EXCHANGE=0),
# focal stage at N = 8 (BASELINE configs[3]: one block sub-encoder per GPU), peer-exchange tests at world size 2.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
run() {  # name, nproc, extra args...
  local name=$1 n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 200 --warmup 10 --no-cpu-baseline --no-operator-api "$@" \
      > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    b = json.loads(open("$OUT/${TAG}_${name}.json").read().strip().splitlines()[-1])
    print("  ", b["n_gpus"], "GPUs", round(b["ms_per_step"], 3), "ms/step", round(b["value"]), "rays/s, e2e", round(b["e2e"]["value"]),
          b["config"].get("exchange", "")[:40], [(k["kernel"], k["ms"]) for k in b["kernels"] if k["kernel"] in ("exchange", "octree_vote", "allreduce")])
except Exception as e:
    print("   unreadable:", e)
PY
}
run bench_n8 8
run bench_n4 4
run bench_n2 2
run bench_n1 1
[ "${NCCL_TOO:-0}" = 1 ] && GF_PEER_EXCHANGE=0 run bench_n8_nccl 8
run bench_focal_n8 8 --workload focal
timeout 300 python -m pytest tests/test_peer_gpu.py -x -q -p no:cacheprovider -s > $OUT/${TAG}_peer_tests.log 2>&1
echo "peer tests rc=$?"; tail -6 $OUT/${TAG}_peer_tests.log
