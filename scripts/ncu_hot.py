"""Top stall sites of a kernel from an ncu report's source page (SASS granularity, with neighbours for context).
   python scripts/ncu_hot.py gpurun_out/prof.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2]
ix = {h: i for i, h in enumerate(hdr)}
S = ix["# Samples"]; SRC = ix["Source"]; EX = ix["Instructions Executed"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
tot = sum(int(r[S] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:top]
for i in order:
    r = body[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols if c < len(r)), reverse=True)[:2]
    print(f"{100*int(r[S] or 0)/tot:5.1f}%  line {i:5d} exec {r[EX]:>8s}  {r[SRC][:90]:90s} {st}")
    for j in range(max(0, i - 3), i):
        print(f"            {j:5d}      {body[j][SRC][:100]}")
