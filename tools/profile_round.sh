#!/bin/bash
# One GPU-box pass that produces everything profiles/ is built from (run through gpurun from the repo root):
#   tools/profile_round.sh <tag>     e.g. r01_v10
# Outputs land in gpurun_out/ (scratch); copy the summaries you want to keep into profiles/.
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $OUT/${TAG}_pytest_gpu.txt
cat $OUT/${TAG}_pytest_gpu.txt
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_case.py > $OUT/${TAG}_memcheck.log 2>&1
timeout 400 compute-sanitizer --tool racecheck python tools/sanitize_case.py > $OUT/${TAG}_racecheck.log 2>&1
grep -H "ERROR SUMMARY\|RACECHECK SUMMARY" $OUT/${TAG}_memcheck.log $OUT/${TAG}_racecheck.log
timeout 600 python bench.py 2> $OUT/${TAG}_bench_native.err | tail -1 > $OUT/${TAG}_bench_native.json
timeout 600 python bench.py --impl reference 2> $OUT/${TAG}_bench_reference.err | tail -1 > $OUT/${TAG}_bench_reference.json
grep "bench\]" $OUT/${TAG}_bench_native.err | tail -12
grep "bench\]" $OUT/${TAG}_bench_reference.err | tail -3
# launch list of the same command (per-launch times are cold-cache and serialised: only the SHARES matter)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-roofline > $OUT/${TAG}_launches.log 2>&1
# one --set full capture of every kernel of one step (for DRAM traffic per launch and the summaries)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gh_" -s 40 -c 8 -o $OUT/${TAG}_full \
    python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu-baseline --no-roofline > $OUT/${TAG}_full.log 2>&1
ls -la $OUT/${TAG}_*
