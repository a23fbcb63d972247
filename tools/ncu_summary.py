"""Markdown + JSON summary of ncu reports: python tools/ncu_summary.py OUT.md OUT.json REP [REP ...]
One row per captured kernel: duration, DRAM bytes, throughput and the issue / shared-memory / occupancy figures the
round's design decisions were based on, plus the top stall reasons."""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue_pct"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefront_pct"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp_insts")]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        yield dict(zip(hdr, vals)), dict(zip(hdr, units))


def to_bytes(v, unit):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    f = float(v)
    return f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def main():
    md_path, js_path, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    md = ["| report | kernel | us | DRAM read MB | DRAM write MB | DRAM % | issue % | smem wavefronts % (conflict share) | warps active % | regs | top stalls (warps per issue) |",
          "|---|---|---|---|---|---|---|---|---|---|---|"]
    js = {}
    for rep in reps:
        name = os.path.basename(rep).replace(".ncu-rep", "")
        for d, u in rows_of(rep):
            k = d.get("Kernel Name", "?")
            short = k.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            if len(short) > 70:
                short = short[:67] + "..."
            g = {}
            for key, nick in KEYS:
                if key in d and d[key] != "":
                    try:
                        g[nick] = float(d[key])
                    except ValueError:
                        pass
            t_us = to_us(d["gpu__time_duration.sum"], u["gpu__time_duration.sum"])
            rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
            wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
            stalls = []
            for key, v in d.items():
                if "issue_stalled" in key and key.endswith("per_issue_active.ratio") and "not_issued" not in key:
                    try:
                        stalls.append((float(v), key.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            top = ", ".join(f"{n} {v:.2f}" for v, n in stalls[:4] if n != "selected")
            conf = g.get("smem_bank_conflicts", 0) / max(g.get("smem_wavefronts", 1), 1)
            md.append(f"| {name} | `{short}` | {t_us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {g.get('dram_pct', 0):.1f} | "
                      f"{g.get('issue_pct', 0):.1f} | {g.get('smem_wavefront_pct', 0):.1f} ({100 * conf:.0f} %) | "
                      f"{g.get('warps_active_pct', 0):.1f} | {int(g.get('regs', 0))} | {top} |")
            js.setdefault(name, []).append({"kernel": k, "us": round(t_us, 2), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
                                            "dram_bytes": int(rd + wr), "issue_pct": g.get("issue_pct"),
                                            "smem_wavefront_pct": g.get("smem_wavefront_pct"),
                                            "smem_bank_conflict_share": round(conf, 3), "warps_active_pct": g.get("warps_active_pct"),
                                            "regs": g.get("regs"), "stalls": {n: round(v, 3) for v, n in stalls[:6]}})
    open(md_path, "w").write("\n".join(md) + "\n")
    json.dump(js, open(js_path, "w"), indent=1)


if __name__ == "__main__":
    main()
