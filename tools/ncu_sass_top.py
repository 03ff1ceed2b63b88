"""Summarise an `ncu --page source --csv` export (SASS view): top instructions by stall samples, samples by opcode class,
and contiguous hot regions.  usage: python tools/ncu_sass_top.py file.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
ins = []
for r in rows[hdr + 1:]:
    if len(r) < len(h): continue
    s = int(r[ix["# Samples"]] or 0)
    ex = int(r[ix["Instructions Executed"]] or 0)
    st = {n: int(r[ix[n]] or 0) for n in stalls}
    ins.append((len(ins), r[ix["Source"]].strip(), s, ex, st))
tot = sum(i[2] for i in ins)
print(f"{len(ins)} SASS instructions, {tot} samples, {sum(i[3] for i in ins)} warp-instructions executed")
agg = collections.Counter()
for i in ins:
    for n, v in i[4].items(): agg[n] += v
print("stall totals:", ", ".join(f"{n[6:]} {v} ({100*v/tot:.0f}%)" for n, v in agg.most_common(9)))
print(f"\ntop {top} instructions:")
for i in sorted(ins, key=lambda x: -x[2])[:top]:
    st = sorted(i[4].items(), key=lambda x: -x[1])[:2]
    print(f"  #{i[0]:5d} {i[2]:6d} {100*i[2]/tot:5.1f}% ex={i[3]:9d}  {i[1][:70]:70s} {[(n[6:], v) for n, v in st if v]}")
# hot regions: windows of 64 instructions
print("\nregions (64-instruction windows) with >= 2% of samples:")
W = 64
for b in range(0, len(ins), W):
    s = sum(i[2] for i in ins[b:b + W])
    if s >= 0.02 * tot:
        ex = sum(i[3] for i in ins[b:b + W])
        ops = collections.Counter(i[1].split()[0] if not i[1].startswith("@") else i[1].split()[1] for i in ins[b:b + W] if i[1])
        print(f"  [{b:5d},{b + W:5d}) {s:6d} {100*s/tot:5.1f}% ex={ex:10d}  {dict(ops.most_common(6))}")
