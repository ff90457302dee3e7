"""Per-source-line sample table from an ncu report (source page, cuda,sass view).
Usage: python tools/ncu_lines.py report.ncu-rep [launch_index] [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # sections start with a "Function Name" row of the kernel; the first header after it belongs to the main file
    launches, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Function Name":
            cur = {"lines": {}}
            launches.append(cur)
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if cur is None or hdr is None or len(r) < len(hdr) or r[2] != "-":
            continue
        try:
            s = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        st = {h[6:]: int(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit()}
        key = (int(r[0]), r[1].strip()[:100])
        e = cur["lines"].setdefault(key, [0, {}, 0])
        e[0] += s
        e[2] += int(r[hdr.index("Instructions Executed")] or 0)
        for k, v in st.items():
            e[1][k] = e[1].get(k, 0) + v
    L = launches[which]["lines"]
    tot = sum(e[0] for e in L.values())
    agg = {}
    for e in L.values():
        for k, v in e[1].items():
            agg[k] = agg.get(k, 0) + v
    print(f"launch {which} of {len(launches)}: {tot} samples; stalls:",
          ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
    for (ln, src), e in sorted(L.items(), key=lambda x: -x[1][0])[:top]:
        st = ", ".join(f"{k} {100 * v / max(e[0], 1):.0f}" for k, v in sorted(e[1].items(), key=lambda x: -x[1])[:3])
        print(f"{100 * e[0] / tot:5.1f}% {e[2]:>11} L{ln:<5} {src[:90]:90s} | {st}")


if __name__ == "__main__":
    main()
