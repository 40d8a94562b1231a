"""Top stall-sample SASS lines of an .ncu-rep (source page)."""
import csv
import subprocess
import sys


def main(path, top=28):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    # first line is the kernel name row
    rows = list(csv.reader(lines[1:]))
    h = rows[0]
    idx = {n: i for i, n in enumerate(h)}
    data = []
    for r in rows[1:]:
        try:
            s = int(r[idx["# Samples"]])
        except Exception:
            continue
        stalls = {k: int(r[idx[k]] or 0) for k in h if k.startswith("stall_") and "Not Issued" not in k and r[idx[k]].isdigit()}
        data.append((s, r[idx["Address"]], r[idx["Source"]], r[idx["Instructions Executed"]], stalls))
    tot = sum(d[0] for d in data)
    print("== %s total samples %d" % (path, tot))
    for s, addr, src, ie, st in sorted(data, key=lambda d: -d[0])[:top]:
        main_st = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print("  %5.1f%%  %-70s exec=%-9s %s" % (100.0 * s / max(tot, 1), src[:70], ie, main_st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 28)
