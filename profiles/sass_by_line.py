"""Attribute ncu per-SASS-instruction counts to CUDA source lines.

    python profiles/sass_by_line.py <report.ncu-rep> <kernel-substring> <lib.so> [launch-index]

ncu's `--page source --csv` is SASS-level without line numbers; `nvdisasm -g` of the same cubin
carries `//## File "...", line N` markers.  Both list the kernel's instructions in address
order, so they are zipped one to one.  Prints instructions executed / stall samples per source
line (inlined call sites are attributed to the innermost line nvdisasm reports).
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def ncu_rows(report, kernel_sub, index=0):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
    sel = [b for b in blocks if kernel_sub in b["name"]]
    return sel[index]


def sass_lines(lib, kernel_sub, want_len=None):
    """Per-instruction (file, line) of the kernel; with several matching .text sections (template
    instantiations) the one whose length equals the ncu row count is used."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    sections = []
    for cubin in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        if kernel_sub not in txt:
            continue
        cur_line, cur = None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur = [] if kernel_sub in m.group(1) else None
                if cur is not None:
                    sections.append(cur)
                continue
            if cur is None:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                cur.append((cur_line, ln.split("*/", 1)[1].strip().rstrip(";")))
    sections = [s for s in sections if s]
    if not sections:
        raise SystemExit("kernel not found in " + lib)
    if want_len is not None:
        for s in sections:
            if len(s) == want_len:
                return s
    return sections[0]


def main():
    report, ksub, lib = sys.argv[1:4]
    index = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    blk = ncu_rows(report, ksub, index)
    ix = {h: i for i, h in enumerate(blk["hdr"])}
    rows = blk["rows"]
    sass = sass_lines(lib, ksub, len(rows))
    print(f"# {blk['name'][:100]}\n# ncu rows {len(rows)}  nvdisasm instrs {len(sass)}")
    n = min(len(rows), len(sass))
    per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    total = 0
    for i in range(n):
        ex = int(rows[i][ix["Instructions Executed"]])
        sm = int(rows[i][ix["# Samples"]])
        key = sass[i][0]
        per[key][0] += ex
        per[key][1] += sm
        per[key][2][rows[i][ix["Source"]].split()[0 if not rows[i][ix["Source"]].strip().startswith("@") else 1].split(".")[0]] += ex
        total += ex
    tot_s = sum(v[1] for v in per.values()) or 1
    print(f"# total warp instructions {total}")
    for key, (ex, sm, ops) in sorted(per.items(), key=lambda kv: -kv[1][0])[:45]:
        top = ", ".join(f"{o}:{c * 100 // max(ex, 1)}%" for o, c in ops.most_common(4))
        print(f"{str(key):38s} instr {100 * ex / total:5.1f}%  samples {100 * sm / tot_s:5.1f}%   {top}")


if __name__ == "__main__":
    main()
