#!/bin/bash
# Last GPU call of a round on one B200: the full parity suite, smoke(), the driver's bench command and a long run, the
# reference arm, the other workloads, the reference's kernels beside ours, then the ncu launch list and one full-set
# capture of every kernel of a step (only after the un-profiled runs have exited 0).
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_end.sh r02at'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/${TAG}_steps.log
echo "== 1. GPU parity tests" | tee $S
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a $S; tail -2 $OUT/${TAG}_gpu_tests.log | tee -a $S
echo "== 2. smoke()" | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1
echo "smoke rc=$?" | tee -a $S; tail -1 $OUT/${TAG}_smoke.log | tee -a $S
echo "== 3. benches (not under a profiler)" | tee -a $S
line() {
python - <<PY | tee -a $S
import json
try:
    b = json.loads(open("$1").read().strip().splitlines()[-1])
    print("  $1: ms/step", round(b["ms_per_step"], 3), b["unit"], round(b["value"]), "e2e", b.get("e2e", {}).get("value"),
          [(k["kernel"], k["ms"]) for k in b.get("kernels", [])[:6]])
except Exception as e:
    print("  $1 unreadable:", e)
PY
}
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench_driver.json 2> $OUT/${TAG}_bench_driver.err
echo "driver-style bench rc=$?" | tee -a $S; line $OUT/${TAG}_bench_driver.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
echo "reference arm rc=$?" | tee -a $S; tail -c 600 $OUT/${TAG}_bench_ref.json | tee -a $S; echo | tee -a $S
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "200-step bench rc=$?" | tee -a $S; line $OUT/${TAG}_bench_n1.json
timeout 600 python bench.py --steps 100 --warmup 10 --hidden 128 --no-cpu-baseline --no-operator-api > $OUT/${TAG}_bench_h128.json 2> $OUT/${TAG}_bench_h128.err
echo "H=128 bench rc=$?" | tee -a $S; line $OUT/${TAG}_bench_h128.json
timeout 600 python bench.py --workload render --steps 10 --warmup 3 --no-cpu-baseline --no-operator-api > $OUT/${TAG}_bench_render.json 2> $OUT/${TAG}_bench_render.err
echo "render bench rc=$?" | tee -a $S; line $OUT/${TAG}_bench_render.json
timeout 600 python bench.py --workload focal --steps 50 --warmup 10 --no-cpu-baseline --no-operator-api > $OUT/${TAG}_bench_focal.json 2> $OUT/${TAG}_bench_focal.err
echo "focal bench rc=$?" | tee -a $S; line $OUT/${TAG}_bench_focal.json
echo "== 4. the reference's kernels beside ours" | tee -a $S
timeout 600 python tools/ref_kernels_timing.py > $OUT/${TAG}_ref_kernels_timing.json 2> $OUT/${TAG}_ref_kernels_timing.err
echo "ref kernels rc=$?" | tee -a $S; tail -c 900 $OUT/${TAG}_ref_kernels_timing.json | tee -a $S; echo | tee -a $S
echo "== 5. ncu launch list (serialised; shares only)" | tee -a $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-operator-api --no-sample-ahead \
    > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a $S
echo "== 6. ncu full-set capture of one step (every kernel)" | tee -a $S
# 13 matched launches per step (sampler, compact, hash fwd, ray bias, MLP fwd, composite fwd / bwd, MLP bwd, ray-bias
# bwd, hash bwd, 3 x Adam): the 4th step
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k "regex:sample_rays|compact_kernel|hash_fwd|hash_bwd|mlp_tc|composite|adam_kernel|ray_bias" \
    --launch-skip ${NCU_SKIP:-39} -c ${NCU_COUNT:-13} -o $OUT/prof_${TAG} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-operator-api --no-sample-ahead \
    > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $S
