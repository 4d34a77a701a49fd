#!/usr/bin/env python
"""Thread-instructions one consumer thread issues per pipeline stage of a K1 (ce_tma_kernel) instantiation, read from
the SASS of the built library — no GPU needed.

    python scripts/sass_fastpath.py [lib.so] <mangled template args, e.g. IfLi7ELi4ELb0ELb1ELb1ELb1E> ...

The walker starts at the head of the innermost loop that contains the max reduction (the consumer loop) and follows
the common path: spin-wait branches fall through, forward branches are taken when they stay inside the loop and skip no
max / exp work (the NaN-aware argmax, the counter flush, the label-dtype variant that is not running), the loop's
back edge ends the walk.  It is a heuristic: use it to compare two builds of the same kernel, and divide by the
pixels per thread (VECP) for the per-pixel figures quoted in DESIGN.md §3.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_instructions(lib, key):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout.splitlines()
    ins, on = [], False
    for line in sass:
        if "Function :" in line:
            on = ("ce_tma_kernel" + key) in line
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def fast_path(ins):
    addr = {a: i for i, (a, _) in enumerate(ins)}
    is_work = lambda t: "MNMX" in t or "MUFU.EX2" in t
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.match(r"BRA (0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
            head = addr[int(m.group(1), 16)]
            if any("MNMX" in ins[j][1] for j in range(head, i)):
                loops.append((head, i))
    head, tail = min(loops, key=lambda c: c[1] - c[0])
    pc, n, ops = head, 0, {}
    while pc <= tail and n < 20000:
        _, t = ins[pc]
        n += 1
        op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
        m = re.search(r"BRA\S* (?:[!UPR\d]+, )*(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) in addr:
            ti = addr[int(m.group(1), 16)]
            before_target = t.split("BRA", 1)[1].split("0x")[0]
            cond = t.startswith("@") or "P" in before_target or "UR" in before_target
            if ti <= pc:
                if not cond:
                    break                                   # the loop's back edge
            elif not cond or (ti <= tail and not any(is_work(ins[j][1]) for j in range(pc + 1, ti))):
                pc = ti
                continue
        pc += 1
    return n, ops


if __name__ == "__main__":
    args = sys.argv[1:]
    lib = os.path.join(ROOT, "cvcs_b200", "libcvcs_b200.so")
    if args and args[0].endswith(".so"):
        lib, args = args[0], args[1:]
    for key in args or ["IfLi7ELi4ELb0ELb1ELb1ELb1E", "I13__nv_bfloat16Li7ELi4ELb0ELb1ELb1ELb1E", "IfLi7ELi4ELb0ELb1ELb0ELb1E",
                        "IfLi7ELi4ELb0ELb1ELb0ELb0E"]:
        ins = kernel_instructions(lib, key)
        if not ins:
            print(f"{key}: no such instantiation in {lib}")
            continue
        n, ops = fast_path(ins)
        top = ", ".join(f"{k} {v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:10])
        print(f"{key}: {n} instructions per thread per stage  ({top})")
