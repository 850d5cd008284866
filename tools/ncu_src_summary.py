"""Summarise `ncu --page source --csv` output: instruction mix and stall samples by opcode class / region."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hdr_i + 1:]
tot_s = sum(int(r[col["# Samples"]] or 0) for r in data)
tot_i = sum(int(r[col["Instructions Executed"]] or 0) for r in data)
by_op = collections.Counter(); by_op_s = collections.Counter()
for r in data:
    op = r[col["Source"]].split()[0] if r[col["Source"]] else "?"
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    op = op.split(".")[0]
    by_op[op] += int(r[col["Instructions Executed"]] or 0)
    by_op_s[op] += int(r[col["# Samples"]] or 0)
print(f"SASS lines {len(data)}  warp-instr {tot_i}  samples {tot_s}")
print("op        instr%   samples%")
for op, n in by_op.most_common(18):
    print(f"{op:9s} {100*n/tot_i:6.1f}  {100*by_op_s[op]/max(tot_s,1):6.1f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[col[h]] or 0) for r in data) for h in stalls}
print("stall samples:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
# hottest instructions
hot = sorted(data, key=lambda r: -int(r[col["# Samples"]] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]
for r in hot:
    print(f"{r[col['Address']]:>8s} {int(r[col['# Samples']] or 0):6d}  {r[col['Source']][:90]}")
