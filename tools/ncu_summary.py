"""Summarise ncu outputs into small tracked files under profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py launches gpurun_out/launches_v3.csv profiles/r1_ncu_launches_v3_summary.csv --passes 3
  python tools/ncu_summary.py full gpurun_out/r1v3_*.ncu-rep profiles/r1_ncu_full_v3.csv
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_write.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:90]


def launches(src, dst, passes):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        v = float(r[iv].replace(",", ""))
        us = v / 1e3 if r[iu] in ("ns", "nsecond") else (v if r[iu] in ("us", "usecond") else v * 1e3)
        a = agg[short(r[ik])]
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        cmd = sys.argv[sys.argv.index("--cmd") + 1] if "--cmd" in sys.argv else f"python tools/prof_step.py {passes}"
        f.write(f"# ncu launch list summary: `ncu --metrics gpu__time_duration.sum --clock-control none {cmd}`\n")
        f.write(f"# {passes} passes (B=16 clouds x 16384 points), {len(rows)} launches, {tot / 1e3:.3f} ms summed "
                f"= {tot / 1e3 / passes:.3f} ms per pass\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("kernel,launches_per_pass,us_per_pass,share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{a[0] / passes:.1f},{a[1] / passes:.1f},{a[1] / tot:.4f}\n")
    print(open(dst).read()[:3000])


def full(reports, dst):
    out = []
    for rep in reports:
        if rep.endswith(".csv"):   # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`)
            txt = open(rep).read()
        else:
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
                 "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            rec = {"report": rep.split("/")[-1], "kernel": short(r[hdr.index("Kernel Name")]),
                   "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
            for k in KEEP:
                if k in hdr:
                    j = hdr.index(k)
                    v = r[j].replace(",", "")
                    if units[j] in scale and v:
                        v = repr(float(v) * scale[units[j]])  # normalised to ns / bytes
                    rec[k] = v
            out.append(rec)
    cols = ["report", "kernel", "grid", "block"] + [k for k in KEEP if any(k in o for o in out)]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, one row per captured launch (units as reported by ncu: time ns, bytes)\n")
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for o in out:
            w.writerow({c: o.get(c, "") for c in cols})
    for o in out:
        t = float(o.get("gpu__time_duration.sum", "0").replace(",", "") or 0)
        rd = float(o.get("dram__bytes_read.sum", "0").replace(",", "") or 0)
        wr = float(o.get("dram__bytes_write.sum", "0").replace(",", "") or 0)
        print(f"{o['kernel'][:60]:60s} grid {o['grid']:>14s} t={t / 1e3:9.1f} us  dram rd {rd / 1e6:8.2f} MB wr {wr / 1e6:8.2f} MB  "
              f"-> {(rd + wr) / max(t, 1):7.1f} GB/s  dram% {o.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', '')}  "
              f"tensor% {o.get('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', '')}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        passes = int(sys.argv[sys.argv.index("--passes") + 1]) if "--passes" in sys.argv else 1
        launches(sys.argv[2], sys.argv[3], passes)
    else:
        full(sys.argv[2:-1], sys.argv[-1])
