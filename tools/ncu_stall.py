"""Source lines ranked by one stall reason: python tools/ncu_stall.py REP KERNEL stall_long_sb [topN]"""
import csv, io, subprocess, sys
rep, kern, col = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; recs = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "Function Name") and len(r) > 8:
        d = {}
        for k, v in zip(hdr, r):
            if k not in d: d[k] = v
        d["file"] = cur; recs.append(d)
def num(d, k):
    try: return float(d.get(k, "0") or 0)
    except ValueError: return 0.0
tot = sum(num(d, col) for d in recs) or 1
recs.sort(key=lambda d: -num(d, col))
print(f"{col}: total {tot:.0f}")
for d in recs[:top]:
    print(f"{num(d, col) / tot * 100:5.1f}%  {d['file'].split('/')[-1]}:{d['Line No']:>4}  {d['Source'].strip()[:120]}")
