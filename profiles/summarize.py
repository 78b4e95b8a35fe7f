"""Turn ncu artifacts brought back in gpurun_out/ into the committed summaries of one round.

    python profiles/summarize.py r01 [tag]      # tag selects gpurun_out/<round>_*_<tag>.*  (default: final)

Writes under profiles/<round>/:
  launches_<tag>.csv / launches_<tag>_summary.txt   per-kernel device time of one run of the hot
                                                    path (ncu gpu__time_duration.sum; cold-cache and
                                                    serialised: compare SHARES, not absolutes)
  <kernel>_<tag>_details.txt                        `ncu --set full` details page
  <kernel>_<tag>_raw.csv                            the raw metrics the roofline uses (DRAM bytes,
                                                    duration, issue activity, registers, occupancy)
  <kernel>_<tag>_by_line.txt                        instructions / stall samples per CUDA source line
"""
import collections
import csv
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(src, dst_dir, tag):
    shutil.copy(src, os.path.join(dst_dir, f"launches_{tag}.csv"))
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        ns = float(row["Metric Value"].replace(",", ""))
        if row.get("Metric Unit", "ns") in ("us", "usecond"):
            ns *= 1e3
        agg.setdefault(name, []).append(ns)
    ours = {k: v for k, v in agg.items() if "avfe" in k or "lm::" in k or k.startswith("void avfe")}
    total = sum(sum(v) for v in ours.values()) or 1.0
    with open(os.path.join(dst_dir, f"launches_{tag}_summary.txt"), "w") as f:
        f.write("libavfe kernels of one `profiles/prof_kernels.py all` run under ncu (time in us; share of libavfe time)\n")
        for k, v in ours.items():
            f.write(f"{k[:70]:70s} launches={len(v):3d} avg_us={sum(v) / len(v) / 1e3:9.1f} share={100 * sum(v) / total:5.1f}%\n")


def full(rep, kernel_key, dst_dir, tag, lib):
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open(os.path.join(dst_dir, f"{kernel_key}_{tag}_details.txt"), "w").write(det)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if rows:
        hdr = rows[0]
        idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
        with open(os.path.join(dst_dir, f"{kernel_key}_{tag}_raw.csv"), "w") as f:
            f.write(",".join(w for w, _ in idx) + "\n")
            for r in rows[1:]:
                f.write(",".join(r[i].replace(",", "") for _, i in idx) + "\n")
    names = sorted({r[hdr.index("Kernel Name")].split("(")[0] for r in rows[2:]}) if rows else []
    for n in names:
        key = n.split("::")[-1].split("<")[0].replace("void ", "").strip()
        out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "sass_by_line.py"), rep, key, lib],
                             capture_output=True, text=True).stdout
        open(os.path.join(dst_dir, f"{key}_{tag}_by_line.txt"), "w").write(out)


def main():
    rnd = sys.argv[1]
    tag = sys.argv[2] if len(sys.argv) > 2 else "final"
    src = os.path.join(ROOT, "gpurun_out")
    dst = os.path.join(ROOT, "profiles", rnd)
    os.makedirs(dst, exist_ok=True)
    lib = os.path.join(ROOT, "avsl_b200", "lib", "libavfe.so")
    lc = os.path.join(src, f"{rnd}_launches_{tag}.csv")
    if os.path.exists(lc):
        launches(lc, dst, tag)
    for key in ("lip", "logmel", "fuse", "fuseln", "fuselntma", "proj", "noise", "logfbank"):
        rep = os.path.join(src, f"{rnd}_{key}_{tag}.ncu-rep")
        if os.path.exists(rep):
            full(rep, key, dst, tag, lib)
    # bench lines are copied by hand from the run they belong to (gpurun_out/ keeps files of earlier
    # rounds under the same names)


if __name__ == "__main__":
    main()
