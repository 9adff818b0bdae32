"""Attribute the samples of an `ncu --page source --csv` (SASS) export to CUDA source lines, using nvdisasm line info of the
object file the kernel was built from:  python tools/ncu_lines.py gpurun_out/ncu_X_src.csv build/attn_bwd_tc.o kernel_substr [top]"""
import csv, re, subprocess, sys, tempfile, os, collections, glob
src_csv, obj, kname = sys.argv[1:4]; top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
cubin = glob.glob(os.path.join(d, "*.cubin"))[0]
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
# walk: find the .text section of the kernel, track "//## File "...", line N" comments, map instruction offset -> line
line_of, cur, infunc = {}, None, False
for ln in txt.splitlines():
    if ln.startswith(".section") or ln.strip().startswith(".section"):
        infunc = (".text." in ln and kname in ln)
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        # inlined-at chains: keep the outermost non-header line if present
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m and cur: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
base = int(data[0][col["Address"]], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
tot = 0.0
for r in data:
    off = int(r[col["Address"]], 16) - base
    key = line_of.get(off, ("?", 0))
    s = float(r[col["# Samples"]] or 0); tot += s
    a = agg[key]; a[0] += s; a[1] += float(r[col["Instructions Executed"]] or 0)
    for st in stalls: a[2][st[6:]] += float(r[col[st]] or 0)
srcs = {}
print(f"{tot:.0f} samples; mapped {sum(v[0] for k, v in agg.items() if k[0] != '?'):.0f}")
for (f, l), (s, x, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", f)) + glob.glob(os.path.join("open-musiclm_b200/csrc", f))
        srcs[f] = open(cand[0]).read().splitlines() if cand else []
    text = srcs[f][l - 1].strip()[:90] if srcs[f] and 0 < l <= len(srcs[f]) else ""
    top2 = ", ".join(f"{k}:{v:.0f}" for k, v in st.most_common(2))
    print(f"{s:7.0f} {100*s/tot:5.1f}%  x{x:10.0f}  {f}:{l:<4d} {text:90s} [{top2}]")
