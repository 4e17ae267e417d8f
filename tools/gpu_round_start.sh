#!/bin/bash
# First GPU call of a round: confirm parity, re-measure, re-profile -- everything the previous round could not
# run once its GPU minutes were spent.  Run through gpurun from the repo root, after `python -c "import
# __graft_entry__ as g; g.build()"` here (which also builds oracle/_ref incl. the nvcc build of the reference's
# own kernels):
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_start.sh r02a'
#
# Outputs land in gpurun_out/ (merged back); turn them into profiles/ with tools/ncu_summary.py.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
echo "== 1. GPU parity tests" | tee $OUT/${TAG}_steps.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/${TAG}_steps.log; tail -3 $OUT/${TAG}_gpu_tests.log | tee -a $OUT/${TAG}_steps.log
echo "== 2. ours beside the reference's kernels compiled by nvcc (opt-in test)" | tee -a $OUT/${TAG}_steps.log
GF_REF_CUDA=1 timeout 300 python -m pytest tests/test_ref_kernels.py -m gpu -q -s -p no:cacheprovider \
    -k nvcc > $OUT/${TAG}_ref_cuda.log 2>&1
echo "ref_cuda rc=$?" | tee -a $OUT/${TAG}_steps.log; grep -E "bit-equal|identical|max abs|passed|failed" $OUT/${TAG}_ref_cuda.log | tee -a $OUT/${TAG}_steps.log
echo "== 3. bench N=1 (not under a profiler)" | tee -a $OUT/${TAG}_steps.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench rc=$?" | tee -a $OUT/${TAG}_steps.log
python - <<PY | tee -a $OUT/${TAG}_steps.log
import json
try:
    b = json.loads(open("$OUT/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
    print("ms/step", round(b["ms_per_step"], 3), "rays/s", round(b["value"]), "e2e", round(b["e2e"]["value"]))
    for k in b.get("kernels", [])[:8]:
        print("  ", k["kernel"], k["ms"], k.get("frac"))
except Exception as e:
    print("bench line unreadable:", e)
PY
echo "== 4. ncu launch list (serialised; shares only)" | tee -a $OUT/${TAG}_steps.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sample-ahead \
    > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a $OUT/${TAG}_steps.log
echo "== 5. ncu full-set capture of the hash kernels (the ones whose working set changed)" | tee -a $OUT/${TAG}_steps.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:hash_ -c 4 -o $OUT/prof_${TAG}_hash \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sample-ahead > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $OUT/${TAG}_steps.log
