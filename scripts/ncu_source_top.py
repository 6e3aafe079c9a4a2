"""Top stalled SASS instructions of a kernel from `ncu -i X.ncu-rep --page source --csv`. Usage: ncu_source_top.py rep [N]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks[:1]:
    h = b["hdr"]; si, ci = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
    stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(float(r[ci]) for r in b["rows"] if r[ci].replace('.', '').isdigit())
    print(b["name"][:80], "total samples", tot)
    order = sorted(range(len(b["rows"])), key=lambda i: -float(b["rows"][i][ci] or 0))[:n]
    for i in sorted(order):
        r = b["rows"][i]; v = float(r[ci])
        top = sorted(((float(r[c] or 0), h[c]) for c in stall_cols), reverse=True)[:2]
        print(f"{i:5d} {v:7.0f} {100 * v / tot:5.1f}%  {r[si].strip()[:70]:70s} {top[0][1]}={top[0][0]:.0f} {top[1][1]}={top[1][0]:.0f}")
