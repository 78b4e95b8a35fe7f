"""Per-SASS-instruction listing (source line, executed count, stall samples) of one kernel.

    python profiles/sass_dump.py <report.ncu-rep> <kernel-substring> <lib.so> [min_sample_pct] [file-substring]
"""
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from sass_by_line import ncu_rows, sass_lines  # noqa: E402


def main():
    report, ksub, lib = sys.argv[1:4]
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
    fsub = sys.argv[5] if len(sys.argv) > 5 else ""
    blk = ncu_rows(report, ksub, 0)
    ix = {h: i for i, h in enumerate(blk["hdr"])}
    rows = blk["rows"]
    sass = sass_lines(lib, ksub, len(rows))
    tot = sum(int(r[ix["# Samples"]]) for r in rows) or 1
    for i, r in enumerate(rows[:len(sass)]):
        sm = int(r[ix["# Samples"]])
        line = sass[i][0]
        if 100.0 * sm / tot >= min_pct and (not fsub or (line and fsub in line[0])):
            print(f"{i:5d} {str(line):34s} ex {int(r[ix['Instructions Executed']]):9d} samp {100.0 * sm / tot:5.2f}%  {r[ix['Source']][:70]}")


if __name__ == "__main__":
    main()
