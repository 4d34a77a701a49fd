#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: instruction-class mix weighted by executions,
top stall lines.  usage: ncu_source_summary.py prof.source.csv[.gz] [n_pixels]"""
import csv, gzip, sys, collections, re
path = sys.argv[1]
npx = float(sys.argv[2]) if len(sys.argv) > 2 else None
op = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(op))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = rows[hdr_i + 1:]
# several launches / pages in one export: keep the first kernel's rows
end = next((i for i, r in enumerate(data) if not r or r[0] in ("Kernel Name", "Address")), len(data))
data = [r for r in data[:end] if len(r) == len(hdr)]
tot_inst = sum(int(r[col["Instructions Executed"]]) for r in data)
tot_thr = sum(int(r[col["Thread Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[col["# Samples"]]) for r in data)
print(f"kernel: {rows[0][1][:100]}")
print(f"warp instr executed {tot_inst:,}  thread instr {tot_thr:,}  samples {tot_samp:,}")
if npx:
    print(f"thread instr / pixel = {tot_thr / npx:.1f}")
mix = collections.Counter()
samp = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[col["Source"]])
    k = m.group(2) if m else "?"
    mix[k] += int(r[col["Instructions Executed"]])
    samp[k] += int(r[col["# Samples"]])
print("opcode mix (share of warp instr | share of stall samples):")
for k, v in mix.most_common(28):
    print(f"  {k:12s} {100 * v / tot_inst:5.1f}%  {100 * samp[k] / max(tot_samp, 1):5.1f}%")
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(int(r[col[s]]) for r in data) for s in stalls}
print("stall reasons (all samples):", ", ".join(f"{k[6:]} {100 * v / max(tot_samp, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print("top sampled instructions:")
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:14]:
    top = sorted(((int(r[col[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"  {int(r[col['# Samples']]):6d}  {r[col['Source']].strip()[:70]:70s} {top}")
