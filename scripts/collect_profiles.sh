#!/bin/bash
# copy the judged summaries of a gpu_final.sh run from gpurun_out/ into profiles/<tag>/
TAG=${1:-r1c}
mkdir -p profiles/$TAG
cp gpurun_out/launches_default.csv gpurun_out/gpu.txt profiles/$TAG/ 2>/dev/null
cp gpurun_out/bench_default.log profiles/$TAG/bench_default.json
cp gpurun_out/bench_reference.log profiles/$TAG/bench_reference.json
cp gpurun_out/bench_variants.log profiles/$TAG/bench_variants.jsonl
cp gpurun_out/kernel_bench.jsonl profiles/$TAG/ 2>/dev/null
for f in gpurun_out/prof_*.details.txt; do   # NOTE: clean the local gpurun_out/ first — gpurun merges, it does not mirror
  n=$(basename $f .details.txt); n=${n#prof_}
  cp gpurun_out/prof_$n.details.txt gpurun_out/prof_$n.source.csv.gz profiles/$TAG/
  python - "$n" "$TAG" <<'PY'
import csv, sys
n, tag = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(f"gpurun_out/prof_{n}.raw.csv")))
with open(f"profiles/{tag}/prof_{n}.raw_metrics.txt", "w") as f:
    for k, u, v in zip(rows[0], rows[1], rows[2]):
        f.write(f"{k}\t{v}\t{u}\n")
PY
done
du -sh profiles/$TAG
