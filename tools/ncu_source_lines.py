"""Per-CUDA-source-line shares of warp-stall samples and executed instructions from an ncu source-page export
(`ncu -i x.ncu-rep --page source --csv --print-source sass,cuda > x_src.csv`).  Run here, no GPU needed.

    python tools/ncu_source_lines.py gpurun_out/x_src.csv [file-substring] [min-share]
"""
import csv
import sys

csv.field_size_limit(10**9)


def sections(path):
    rows = list(csv.reader(open(path)))
    secs, i = [], 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "File Path":
            secs.append({"file": r[1], "func": rows[i + 1][1], "hdr": rows[i + 2], "rows": []})
            i += 3
            continue
        if secs and r:
            secs[-1]["rows"].append(r)
        i += 1
    return secs


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.012
    for s in sections(path):
        h = s["hdr"]
        if "Warp Stall Sampling (All Samples)" not in h or want not in s["file"]:
            continue
        ia, ie = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        lines = [r for r in s["rows"] if len(r) > ie and r[2] == "-" and r[0].isdigit()]
        tot, ti = sum(int(r[ia]) for r in lines), sum(int(r[ie]) for r in lines)
        if not tot:
            continue
        print(f"===== {s['file'].split('/')[-1]}  {s['func'][:90]}  samples {tot}  warp-instructions {ti}")
        sb, sw = h.index("stall_barrier"), h.index("stall_wait")
        for r in sorted(lines, key=lambda r: int(r[0])):
            if int(r[ia]) > tot * thr or int(r[ie]) > ti * thr:
                st = {h[j]: int(r[j]) for j in range(sb, sw + 1) if r[j].isdigit() and int(r[j]) > 0}
                top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
                print(f"L{r[0]:>4} smp {100 * int(r[ia]) / tot:5.1f}% ins {100 * int(r[ie]) / max(ti, 1):5.1f}%  {r[1].strip()[:88]}  {top}")


if __name__ == "__main__":
    main()
