"""Turn the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/ (run on the CPU box).
    python scripts/summarise_profiles.py r2"""
import csv, io, json, subprocess, sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
out = ROOT / "profiles"

# ---- launch list
src = ROOT / "gpurun_out" / f"{tag}_launches.csv"
if src.exists():
    lines = [l for l in src.read_text().splitlines() if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
    agg = defaultdict(list)
    for r in rows:
        if r["Metric Name"] == "gpu__time_duration.sum":
            agg[r["Kernel Name"]].append(float(r["Metric Value"]) / 1e3)
    total = sum(sum(v) for v in agg.values())
    text = [f"ncu --metrics gpu__time_duration.sum --clock-control none -k regex:<the library's kernels> -c 400, python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-verify "
            f"(cfg4 then cfg2 in one process, 1 GPU); per-launch times are cold-cache and serialised: compare SHARES. {len(rows)} launches, {total / 1e3:.1f} ms"]
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        text.append(f"{name[:110]:110s} n={len(v):4d} mean={sum(v) / len(v):12.1f} us share={100 * sum(v) / total:5.1f}%")
    (out / f"{tag}_launches_summary.txt").write_text("\n".join(text) + "\n")
    (out / f"{tag}_launches.csv").write_text("\n".join(lines) + "\n")
    print("\n".join(text[:12]))

# ---- full captures
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__t_bytes.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum"]
traffic = {}
tp = out / "roofline_traffic.json"
if tp.exists():
    traffic = json.loads(tp.read_text())
for rep in sorted((ROOT / "gpurun_out").glob(f"{tag}_*.ncu-rep")):
    r = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[-1]
    rec = {"report": rep.name, "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None,
           "command": "ncu --set full --clock-control none --import-source on (one launch; profiler replays: timings are not bench values)"}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            try:
                rec[h] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                rec[h] = {"value": v, "unit": u}
    def val(name):
        x = rec.get(name)
        if not x:
            return 0.0
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(x["unit"], 1.0)
        return x["value"] * scale
    rec["dram_bytes_per_launch"] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    (out / (rep.stem + "_ncu_summary.json")).write_text(json.dumps(rec, indent=1) + "\n")
    print(rep.name, rec.get("gpu__time_duration.sum"), "tensor active % elapsed:", rec.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", {}).get("value"),
          "dram MB:", round(rec["dram_bytes_per_launch"] / 1e6, 1))
    if "knn_screen" in (rec["kernel"] or ""):
        cfg = "cfg4" if "cfg4" in rep.name else "cfg2"
        traffic[cfg] = {"dram_bytes_per_launch": int(rec["dram_bytes_per_launch"]),
                        "source": f"profiles/{rep.stem}_ncu_summary.json (dram__bytes_read.sum + dram__bytes_write.sum, one launch of knn_screen_kernel"
                                  + (", 18944 queries x 10 M rows: one of the ~5.3 launches of a cfg4 step)" if cfg == "cfg4" else ")")}
tp.write_text(json.dumps(traffic, indent=1) + "\n")

# ---- SASS mnemonic histogram of the shipped library
so = ROOT / "agplace_b200" / "libagpknn.so"
r = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True)
hist = defaultdict(int)
for line in r.stdout.splitlines():
    line = line.strip()
    if line.startswith("/*") and "*/" in line:
        body = line.split("*/", 1)[1].strip()
        if not body or body.startswith("/*"):
            continue
        tok = body.split()
        if tok and tok[0].startswith("@"):
            tok = tok[1:]
        if tok:
            hist[tok[0].rstrip(";")] += 1
keys = sorted(hist.items(), key=lambda kv: -kv[1])
blackwell = {k: v for k, v in hist.items() if k.startswith(("UTC", "LDTM", "STTM", "UTMA", "UBLKCP", "SYNCS", "UTMAPF", "UCGABAR", "HMMA", "HGMMA"))}
text = ["cuobjdump -sass agplace_b200/libagpknn.so | mnemonic histogram (all kernels, sm_100a)", "", "Blackwell-specific / tensor / TMA mnemonics:"]
text += [f"  {k:32s} {v}" for k, v in sorted(blackwell.items())]
text += ["", "top 40 mnemonics:"] + [f"  {k:32s} {v}" for k, v in keys[:40]]
(out / f"{tag}_sass_mnemonics.txt").write_text("\n".join(text) + "\n")
print("\n".join(text[:16]))
