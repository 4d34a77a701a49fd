mkdir -p gpurun_out; rm -f gpurun_out/variants.log
run() { echo "== $*" >> gpurun_out/variants.log; timeout 100 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-copy-ref "$@" >> gpurun_out/variants.log 2>&1; }
run --workload cfg3
run --workload cfg5
run --workload tile13
run --workload cfg2 --layout nhwc
run --workload cfg2 --metrics-only
run --workload cfg2 --label-dtype i64
run --workload cfg3 --layout nhwc
python - <<'PY'
import json
for l in open('gpurun_out/variants.log'):
    if l.startswith('=='): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('   ', round(d['value'],2), d['unit'], 'frac', round(d['roofline']['frac'],3), 'GB/s', round(d['roofline']['achieved'],1))
PY
