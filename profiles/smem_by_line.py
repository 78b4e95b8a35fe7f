"""Shared-memory wavefronts (total / ideal / excessive) per CUDA source line, from an ncu report.

    python profiles/smem_by_line.py <report.ncu-rep> <kernel-substring> <lib.so> [launch-index]

Same zip of ncu's SASS-level source page with nvdisasm line markers as sass_by_line.py."""
import collections
import sys

from sass_by_line import ncu_rows, sass_lines


def main():
    report, ksub, lib = sys.argv[1:4]
    index = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    blk = ncu_rows(report, ksub, index)
    ix = {h: i for i, h in enumerate(blk["hdr"])}
    rows = blk["rows"]
    sass = sass_lines(lib, ksub, len(rows))
    cols = [c for c in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Wavefronts Shared Excessive") if c in ix]
    per = collections.defaultdict(lambda: [0] * len(cols))
    for i in range(min(len(rows), len(sass))):
        for j, c in enumerate(cols):
            v = rows[i][ix[c]].replace(",", "")
            per[sass[i][0]][j] += int(float(v)) if v not in ("", "-") else 0
    tot = [sum(v[j] for v in per.values()) for j in range(len(cols))]
    print("# " + blk["name"][:100])
    print("# totals: " + ", ".join(f"{c} = {t}" for c, t in zip(cols, tot)))
    for key, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:30]:
        if v[0]:
            print(f"{str(key):38s} " + "  ".join(f"{c.split()[-1] if c != cols[0] else 'total'} {x:10d} ({100 * x / max(tot[0], 1):4.1f}%)" for c, x in zip(cols, v)))


if __name__ == "__main__":
    main()
