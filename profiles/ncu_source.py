#!/usr/bin/env python
"""Per-instruction stall samples of one kernel from an .ncu-rep (source page, SASS view).

    python profiles/ncu_source.py gpurun_out/prof.ncu-rep wr_fwd [min_samples]

Prints every SASS instruction with at least `min_samples` warp-stall samples, with its dominant
stall reasons, plus the total per stall reason."""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    min_samples = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
    hdr = rows[0]
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    i_src, i_samp = hdr.index("Source"), hdr.index("# Samples")
    total = Counter()
    n_all = 0
    for k, r in enumerate(rows[1:]):
        if len(r) != len(hdr):
            continue
        if r[i_samp] == "# Samples":      # a second section (another view of the same kernel)
            break
        samp = int(r[i_samp] or 0)
        n_all += samp
        st = {hdr[i][6:]: int(r[i] or 0) for i in stall_cols}
        for kk, v in st.items():
            total[kk] += v
        if samp >= min_samples:
            top = ", ".join(f"{kk}={v}" for kk, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
            print(f"{k:5d} {samp:6d}  {r[i_src].strip()[:90]:90s} {top}")
    print(f"total samples {n_all}; by reason: " + ", ".join(f"{k}={v}" for k, v in total.most_common(10)))


if __name__ == "__main__":
    main()
