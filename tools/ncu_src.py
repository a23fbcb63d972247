"""Per-source-line hot spots from an ncu report:
   python tools/ncu_src.py REP KERNEL [launch_skip] [topN]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
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
ks = "Warp Stall Sampling (All Samples)"; ki = "Instructions Executed"
tot_s = sum(num(d, ks) for d in recs) or 1; tot_i = sum(num(d, ki) for d in recs) or 1
print(f"total samples {tot_s:.0f}  total warp-instr {tot_i:.0f}")
extra = ["L1 Wavefronts Shared Excessive", "L2 Theoretical Sectors Global Excessive"]
recs.sort(key=lambda d: -num(d, ks))
for d in recs[:top]:
    print(f"{num(d,ks)/tot_s*100:5.1f}% smp {num(d,ki)/tot_i*100:5.1f}% ins  exc_smem {num(d,extra[0]):>10.0f}  {d['file'].split('/')[-1]}:{d['Line No']:>4}  {d['Source'].strip()[:100]}")
