"""Summarise an `ncu --page source --csv` export (SASS view): stall mix, samples per opcode, hottest instructions.
  python tools/ncu_src_show.py file_src.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
S = col["# Samples"]; X = col["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[S] or 0) for r in data); totx = sum(float(r[X] or 0) for r in data)
print(f"{path}: {len(data)} SASS instructions, {tot:.0f} samples, {totx:.3g} warp-instructions executed")
mix = {s: sum(float(r[col[s]] or 0) for r in data) for s in stalls}
print("stall mix:", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:10]))
by_op = collections.Counter(); by_opx = collections.Counter()
for r in data:
    src = r[col["Source"]].strip(); op = src.split()[0] if src else "?"
    if op.startswith("@"): op = src.split()[1]
    op = op.split(".")[0] + ("." + op.split(".")[1] if "." in op and op.split(".")[0] in ("LDS", "STS", "MUFU", "LDG", "STG", "SYNCS", "UTCBAR", "LDTM", "BAR", "RED", "ATOMG") else "")
    by_op[op] += float(r[S] or 0); by_opx[op] += float(r[X] or 0)
print("opcode: samples% | executed%")
for op, v in by_op.most_common(25):
    print(f"  {op:14s} {100*v/tot:5.1f}%  {100*by_opx[op]/totx:5.1f}%")
print("hottest instructions:")
for r in sorted(data, key=lambda r: -float(r[S] or 0))[:top]:
    st = sorted(((float(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"  {r[col['Address']][-5:]} {float(r[S]):7.0f} {100*float(r[S])/tot:5.2f}%  {r[col['Source']][:70]:70s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
