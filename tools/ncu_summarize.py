"""Turn `ncu --page raw --csv` exports (tools/ncu_capture.sh) into the small text/JSON files kept under profiles/.
  python tools/ncu_summarize.py <round-tag> <gemm_raw.csv> [other_raw.csv ...]
Writes profiles/<tag>_ncu_summary.txt (one line per captured launch) and profiles/<tag>_ncu_gemm_traffic.json
(DRAM bytes of the GEMM family over one full training step: read by bench.py for roofline.traffic)."""
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
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("smsp__inst_executed.sum", "warp_inst")]


def short(name):
    return name.replace("void ", "").replace("omlm::", "").split("(")[0]


def to_unit(val, unit, want):
    """ncu prints each metric in its own unit (byte / Kbyte / Mbyte, ns / us / ms): normalise."""
    v = float(val.replace(",", ""))
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    if want in ("MB", "us") and unit in scale:
        return v * scale[unit]
    return v


def rows_of(path):
    raw = list(csv.reader(open(path)))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    ix = {c: hdr.index(c) for c, _ in COLS if c in hdr}
    kn = hdr.index("Kernel Name")
    out = []
    for r in rows:
        d = {"name": short(r[kn])}
        for c, lbl in COLS:
            if c in ix:
                want = "MB" if lbl.endswith("_MB") else ("us" if lbl == "us" else "")
                try:
                    d[lbl] = to_unit(r[ix[c]], units[ix[c]], want)
                except ValueError:
                    d[lbl] = r[ix[c]]
        out.append(d)
    return out


def main():
    tag, gemm_csv, others = sys.argv[1], sys.argv[2], sys.argv[3:]
    lines = [f"# ncu --set full --clock-control none; per-launch values (times under ncu are cold-cache and serialised: use the shares, not the absolutes)",
             "kernel | " + " | ".join(lbl for _, lbl in COLS)]
    fmt = lambda d: d["name"] + " | " + " | ".join(f"{d[lbl]:.6g}" if isinstance(d.get(lbl), float) else str(d.get(lbl, "")) for _, lbl in COLS)
    g = rows_of(gemm_csv)
    gem = [d for d in g if "gemm_bf16_kernel" in d["name"] or "gemm_ffn_up_kernel" in d["name"]]
    lines.append(f"## GEMM family, every launch of one full cfg2 training step ({len(gem)} launches)")
    lines += [fmt(d) for d in gem]
    total_mb = sum(d["rd_MB"] + d["wr_MB"] for d in gem)
    total_us = sum(d["us"] for d in gem)
    by = collections.OrderedDict()
    for d in gem:
        a = by.setdefault(d["name"], [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += d["us"]; a[2] += d["rd_MB"] + d["wr_MB"]; a[3] += d["tensor%"] * d["us"]
    lines.append("")
    lines.append(f"GEMM family over the step: {len(gem)} launches, {total_us:.0f} us under ncu, {total_mb:.0f} MB DRAM = {total_mb / len(gem):.1f} MB per launch")
    for k, (n, us, mb, tw) in by.items():
        lines.append(f"  {k:42s} n={n:3d}  {us:8.1f} us  {mb:8.0f} MB  tensor-pipe active (time-weighted) {tw / us:5.1f}%")
    traffic = dict(source=f"profiles/{tag}_ncu_summary.txt (ncu --set full over every GEMM launch of one cfg2 step, dram__bytes_read.sum + dram__bytes_write.sum)",
                   launches_per_step=len(gem), family_mbytes_per_step=total_mb, bytes_per_launch=total_mb * 1e6 / len(gem), family_us_under_ncu=total_us)
    json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_gemm_traffic.json"), "w"), indent=1)
    for f in others:
        lines.append("")
        lines.append(f"## {os.path.basename(f)}")
        lines += [fmt(d) for d in rows_of(f)]
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-30:]))


if __name__ == "__main__":
    main()
