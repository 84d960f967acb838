"""SM-time budget of one pass of the streamed pipeline's work from an ncu launch list (run here, no GPU needed).

    ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,sm__cycles_active.avg,launch__grid_size,launch__block_size \
        --clock-control none --csv --log-file gpurun_out/<tag>_launches.csv python tools/prof_two_phase.py 3
    python tools/sm_budget.py gpurun_out/<tag>_launches.csv profiles/r2_sm_budget.json profiles/r2_ncu_launches_summary.csv --passes 3 \
        [--step-ms 1.467]

The streamed pipeline overlaps the coordinate phase of five batches with the feature phase of a sixth, so what bounds
a step is not a kernel's duration but the SM-time the step's kernels need in total against the 148 SMs x step time the
machine offers.  Per kernel: SM-time = sum over launches of sm__cycles_active.sum (cycles during which an SM had at
least one warp of the launch resident) / SM clock, i.e. SM x ms.  ncu serialises launches and runs them cold, so
durations are upper bounds; SHARES are what the budget is about.
"""
import csv
import json
import os
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("ws3d::<unnamed>::", "")
    return name[:80]


FAMILY = [("fps", "sampling (FPS)"), ("sa_mlp_fused", "fused SA scale (tcgen05)"), ("mlp_layer", "shared-MLP layer (tcgen05)"),
          ("three_interpolate", "interpolation"), ("three_nn", "three_nn stencils"), ("grid_query", "ball query"),
          ("ball_query", "ball query"), ("grid_build", "cell-grid build"), ("group_affine", "grouping (affine gather)"),
          ("group_", "grouping"), ("split_pointcloud", "split"),
          ("at::", "ATen (one-time weight folding in pass 1, the tool's final reduction)")]


def family(k):
    for key, fam in FAMILY:
        if key in k:
            return fam
    return "other"


def main():
    src, dst_json, dst_csv = sys.argv[1], sys.argv[2], sys.argv[3]
    passes = int(sys.argv[sys.argv.index("--passes") + 1]) if "--passes" in sys.argv else 1
    step_ms = float(sys.argv[sys.argv.index("--step-ms") + 1]) if "--step-ms" in sys.argv else None
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    col = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Unit", "Metric Value")}
    launches = defaultdict(dict)
    for r in rows:
        d = launches[r[col["ID"]]]
        d["kernel"] = short(r[col["Kernel Name"]])
        d["grid"] = r[col["Grid Size"]]
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        name = r[col["Metric Name"]]
        if name == "gpu__time_duration.sum":
            v = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)   # -> us
        d[name] = v
    agg = defaultdict(lambda: {"launches": 0, "us": 0.0, "sm_cycles": 0.0, "grids": set()})
    clock_mhz = []
    for d in launches.values():
        a = agg[d["kernel"]]
        a["launches"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["sm_cycles"] += d.get("sm__cycles_active.sum", 0.0)
        a["grids"].add(d["grid"])
        if d.get("sm__cycles_active.avg") and d.get("gpu__time_duration.sum", 0) > 20:
            # a kernel that keeps its SMs busy for its whole duration: active cycles / duration ~ the SM clock
            clock_mhz.append(d["sm__cycles_active.avg"] / d["gpu__time_duration.sum"])
    # SM clock: the driver-measured maximum (clocks sampled under load sit there); fall back to the busiest kernel's
    # active-cycles / duration ratio
    clk = None
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            clk = float(json.load(f)["sm_max_mhz"])
    except Exception:
        pass
    if not clk:
        clk = min(max(max(clock_mhz) if clock_mhz else 1965.0, 1000.0), 2100.0)
    tot_us = sum(a["us"] for a in agg.values())
    kernels, fams = [], defaultdict(lambda: {"sm_ms": 0.0, "serial_ms": 0.0})
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["sm_cycles"]):
        sm_ms = a["sm_cycles"] / clk / 1e3 / passes          # SM x ms per pass
        kernels.append({"kernel": k, "launches_per_pass": round(a["launches"] / passes, 1), "serial_us_per_pass": round(a["us"] / passes, 1),
                        "sm_ms_per_pass": round(sm_ms, 3), "avg_sms_busy": round(a["sm_cycles"] / clk / max(a["us"], 1e-9), 1),
                        "grids": sorted(a["grids"])[:4]})
        f = fams[family(k)]
        f["sm_ms"] += sm_ms
        f["serial_ms"] += a["us"] / passes / 1e3
    total_sm_ms = sum(f["sm_ms"] for f in fams.values())
    out = {"source": "ncu launch list of tools/prof_two_phase.py (coordinate phase in FPS throughput mode + feature phase, eager, serialised, cold)",
           "passes": passes, "sm_clock_mhz_estimate": round(clk, 1), "serial_ms_per_pass": round(tot_us / passes / 1e3, 3),
           "sm_ms_per_pass": round(total_sm_ms, 2),
           "families": {k: {"sm_ms": round(v["sm_ms"], 2), "share_of_sm_time": round(v["sm_ms"] / total_sm_ms, 4), "serial_ms": round(v["serial_ms"], 3)}
                        for k, v in sorted(fams.items(), key=lambda kv: -kv[1]["sm_ms"])},
           "kernels": kernels[:24]}
    if step_ms:
        offered = 148 * step_ms
        out["streamed_step_ms"] = step_ms
        out["sm_ms_offered_per_step"] = round(offered, 1)
        out["sm_time_needed_over_offered"] = round(total_sm_ms / offered, 3)
        out["note"] = ("needed / offered = the fraction of the machine's SM-time one step's kernels occupy when each runs alone and cold; "
                       "the remainder is what overlap cannot fill (dependent chains of sub-wave launches) plus launch gaps")
    with open(dst_json, "w") as f:
        json.dump(out, f, indent=1)
    with open(dst_csv, "w") as f:
        f.write(f"# ncu launch list summary ({passes} passes, B=16 x 16384): per-launch times are cold-cache and serialised: compare SHARES\n")
        f.write("kernel,launches_per_pass,serial_us_per_pass,share_of_serial,sm_ms_per_pass,share_of_sm_time,avg_sms_busy\n")
        for k in sorted(kernels, key=lambda r: -r["serial_us_per_pass"]):
            f.write(f"\"{k['kernel']}\",{k['launches_per_pass']},{k['serial_us_per_pass']},{k['serial_us_per_pass'] * passes / tot_us:.4f},"
                    f"{k['sm_ms_per_pass']},{k['sm_ms_per_pass'] / total_sm_ms:.4f},{k['avg_sms_busy']}\n")
    print(json.dumps({k: out[k] for k in out if k != "kernels"}, indent=1))
    for k in kernels[:16]:
        print(k)


if __name__ == "__main__":
    main()
