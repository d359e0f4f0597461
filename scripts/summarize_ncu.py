"""Turns gpurun_out/*.ncu-rep and launches_*.csv into small text summaries under profiles/ (run in the build
container: ncu reads reports without a GPU).   python scripts/summarize_ncu.py <tag>"""
import csv, io, os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles"); GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]

def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    return rows[0], rows[1], rows[2:]

def summarize_report(rep, out):
    hdr, units, rows = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {os.path.basename(rep)} (cold-cache, serialised replays)\n")
        for r in rows:
            f.write(f"\n## {r[idx['Kernel Name']][:140]}\n")
            for k in KEYS:
                if k in idx and r[idx[k]] != "":
                    f.write(f"{k:85s} {r[idx[k]]:>18s} {units[idx[k]]}\n")
            rd = float(r[idx["dram__bytes_read.sum"]] or 0); wr = float(r[idx["dram__bytes_write.sum"]] or 0)
            f.write(f"{'traffic = dram read + write':85s} {rd:.3f} {units[idx['dram__bytes_read.sum']]} + {wr:.3f} {units[idx['dram__bytes_write.sum']]}\n")
            vals = sorted(((float(r[idx[k]] or 0), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k in stall), reverse=True)
            tot = sum(v for v, _ in vals) or 1
            f.write("stall samples: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for v, k in vals[:6]) + "\n")

def summarize_launches(path, out):
    txt = open(path).read()
    start = txt.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void seqdiff::", "").strip()
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit in ("ms", "msecond") else 1.0)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values()) or 1
    with open(out, "w") as f:
        f.write(f"# per-kernel device time over {sum(a[0] for a in agg.values())} consecutive launches ({os.path.basename(path)}; ncu "
                "gpu__time_duration.sum, --clock-control none; cold-cache + serialised: compare SHARES)\n")
        f.write(f"{'kernel':90s} {'launches':>8s} {'total us':>10s} {'share':>7s}\n")
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name[:90]:90s} {n:8d} {us:10.1f} {100 * us / tot:6.1f}%\n")

os.makedirs(OUT, exist_ok=True)
for f in sorted(os.listdir(GO)):
    if f.endswith(f"_{tag}.ncu-rep"):
        summarize_report(os.path.join(GO, f), os.path.join(OUT, f.replace(".ncu-rep", ".summary.txt")))
        print("wrote", f.replace(".ncu-rep", ".summary.txt"))
    if f == f"launches_{tag}.csv":
        summarize_launches(os.path.join(GO, f), os.path.join(OUT, f"launches_{tag}.summary.txt"))
        print("wrote launches summary")
