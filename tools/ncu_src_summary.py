"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` output: per kernel, instruction mix and
stall samples by opcode, plus the hottest instructions.  Usage: ncu_src_summary.py file.csv [n_hot]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 12
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for si, hdr_i in enumerate(starts):
    end = starts[si + 1] - 1 if si + 1 < len(starts) else len(rows)
    name = rows[hdr_i - 1][1] if hdr_i > 0 and rows[hdr_i - 1] and rows[hdr_i - 1][0] == "Kernel Name" else "?"
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hdr_i + 1:end] if len(r) == len(hdr)]

    def num(r, h):
        try:
            return int(r[col[h]] or 0)
        except ValueError:
            return 0
    tot_s = sum(num(r, "# Samples") for r in data)
    tot_i = sum(num(r, "Instructions Executed") for r in data)
    by_op, by_op_s = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[col["Source"]].split()
        op = toks[0] if toks else "?"
        if op.startswith("@") and len(toks) > 1:
            op = toks[1]
        op = op.split(".")[0]
        by_op[op] += num(r, "Instructions Executed")
        by_op_s[op] += num(r, "# Samples")
    print(f"=== {name}\nSASS lines {len(data)}  warp-instr {tot_i}  samples {tot_s}")
    print("op        instr%   samples%")
    for op, n in by_op.most_common(18):
        print(f"{op:9s} {100 * n / max(tot_i, 1):6.1f}  {100 * by_op_s[op] / max(tot_s, 1):6.1f}")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(num(r, h) for r in data) for h in stalls}
    print("stall samples:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:nhot]:
        print(f"{r[col['Address']]:>8s} {num(r, '# Samples'):6d}  {r[col['Source']][:90]}")
