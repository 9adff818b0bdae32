"""Turn the CSV exports of tools/ncu_capture.sh into the small text/JSON files kept under profiles/.
  python tools/ncu_summarize.py gpurun_out/ncu_hot_raw.csv [gpurun_out/ncu_src_walk.csv gpurun_out/ncu_src_ffnup.csv]
Writes profiles/r01_ncu_summary.txt and profiles/r01_ncu_gemm_traffic.json (read by bench.py for roofline.traffic)."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd_MB"), ("dram__bytes_write.sum", "wr_MB"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("smsp__inst_executed.sum", "warp_inst")]
LAYERS, GEMMS_PER_LAYER = 6, 15


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0]


def main():
    raw = list(csv.reader(open(sys.argv[1])))
    hdr, rows = raw[0], raw[2:]
    ix = {c: hdr.index(c) for c, _ in COLS if c in hdr}
    kn = hdr.index("Kernel Name")
    out = ["# ncu --set full --clock-control none, one depth-1 coarse training step at the bench shapes (B=16, N=1024, d=1024)",
           "# launch order; per-launch values (times under ncu are cold-cache and serialised: use the shares, not the absolutes)",
           "kernel | " + " | ".join(lbl for c, lbl in COLS if c in ix)]
    gemm = []
    for r in rows:
        vals = []
        for c, lbl in COLS:
            if c in ix:
                v = r[ix[c]]
                try:
                    v = f"{float(v):.6g}"
                except ValueError:
                    pass
                vals.append(v)
        out.append(short(r[kn]) + " | " + " | ".join(vals))
        if "gemm_bf16_kernel" in r[kn] or "gemm_ffn_up_kernel" in r[kn]:
            gemm.append(dict(name=short(r[kn]), us=float(r[ix["gpu__time_duration.sum"]]), grid=int(r[ix["launch__grid_size"]]),
                             mb=float(r[ix["dram__bytes_read.sum"]]) + float(r[ix["dram__bytes_write.sum"]])))
    # GEMM family traffic, scaled from the depth-1 capture to the 6-layer bench step: the per-layer GEMMs are the
    # 15 launches between the rel-pos MLP / logit-head launches (5 forward incl. the fused FFN-up, 10 backward)
    n = len(gemm)
    heads = [g for i, g in enumerate(gemm) if i < 2 or 7 <= i < 16 or i >= n - 4]     # rel-pos fwd, logits fwd+bwd, rel-pos bwd
    layer = [g for i, g in enumerate(gemm) if not (i < 2 or 7 <= i < 16 or i >= n - 4)]
    assert len(layer) == GEMMS_PER_LAYER, (len(layer), n)
    total_mb = LAYERS * sum(g["mb"] for g in layer) + sum(g["mb"] for g in heads)
    launches = LAYERS * len(layer) + len(heads)
    traffic = dict(source="profiles/r01_ncu_hot_raw.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)",
                   launches_per_step=launches, family_mbytes_per_step=total_mb, bytes_per_launch=total_mb * 1e6 / launches,
                   per_layer=[dict(name=g["name"], grid=g["grid"], us=g["us"], mbytes=g["mb"]) for g in layer])
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "r01_ncu_gemm_traffic.json"), "w"), indent=1)
    out.append("")
    out.append(f"GEMM family DRAM traffic scaled to the 6-layer step: {total_mb:.0f} MB over {launches} launches = "
               f"{total_mb / launches:.1f} MB per launch")
    for f in sys.argv[2:]:
        src = list(csv.reader(open(f)))
        h, body = src[1], src[2:]
        isrc, iex, ist = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
        hist, stall = collections.Counter(), collections.Counter()
        for r in body:
            try:
                e, t = int(r[iex] or 0), int(r[ist] or 0)
            except (ValueError, IndexError):
                continue
            p = r[isrc].split()
            op = p[1] if p and p[0].startswith("@") and len(p) > 1 else (p[0] if p else "?")
            op = op.split(".")[0]
            hist[op] += e; stall[op] += t
        tot = sum(hist.values())
        out.append("")
        out.append(f"## SASS opcode mix, {os.path.basename(f)} (warp instructions executed, % of kernel, stall samples)")
        for op, c in hist.most_common(16):
            out.append(f"{op:10s} {c:>12d} {100.0 * c / tot:5.1f}%  {stall[op]}")
    open(os.path.join(ROOT, "profiles", "r01_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[-40:]))


if __name__ == "__main__":
    main()
