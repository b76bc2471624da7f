"""Top SASS lines by warp-stall samples from an `ncu --set full --import-source on` report (source page, CSV).
   python profiles/stall_lines.py gpurun_out/x.ncu-rep [N]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main(path, top=30):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = rows[2:]
    tot = Counter()
    lines = []
    for k, r in enumerate(data):
        if len(r) < len(hdr):
            continue
        n = int(r[ix["# Samples"]] or 0)
        st = {c: int(r[ix[c]] or 0) for c in stall_cols}
        for c, v in st.items():
            tot[c] += v
        lines.append((n, k, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), st))
    total = sum(l[0] for l in lines)
    print("total samples %d; by reason: %s" % (total, tot.most_common(6)))
    for n, k, src, ex, st in sorted(lines, key=lambda l: -l[0])[:top]:
        main_st = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print("%6d samples %5.1f%% line %5d %9d exec  %-70s %s" % (n, 100.0 * n / total, k, ex, src[:70], main_st))
    return lines


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
