"""Split a kernel's SASS (ncu --page source --csv --print-source sass) into runs of equal execution count:
shows where the executed instructions are.  Usage: ncu_sass_regions.py file.csv <kernel-substring> [min_share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for si, h in enumerate(starts):
    name = rows[h - 1][1] if rows[h - 1] and rows[h - 1][0] == "Kernel Name" else "?"
    if pat not in name:
        continue
    end = starts[si + 1] - 1 if si + 1 < len(starts) else len(rows)
    hdr = rows[h]; col = {c: i for i, c in enumerate(hdr)}
    data = [r for r in rows[h + 1:end] if len(r) == len(hdr)]
    ex = [int(r[col["Instructions Executed"]] or 0) for r in data]
    tot = sum(ex)
    print(f"=== {name}: {len(data)} SASS, {tot} warp-instr")
    i = 0
    while i < len(data):
        j = i
        while j + 1 < len(data) and abs(ex[j + 1] - ex[i]) <= 0.02 * max(ex[i], 1):
            j += 1
        share = sum(ex[i:j + 1]) / max(tot, 1)
        if share >= minshare:
            ops = {}
            for r in data[i:j + 1]:
                t = r[col["Source"]].split()
                op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
                ops[op] = ops.get(op, 0) + 1
            top = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:8])
            smp = sum(int(r[col["# Samples"]] or 0) for r in data[i:j + 1])
            print(f"  lines {i:5d}-{j:5d} ({j - i + 1:4d})  exec/line {ex[i]:9d}  share {100 * share:5.1f}%  samples {smp:6d}  {top}")
        i = j + 1
    break
