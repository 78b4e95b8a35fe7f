"""SASS opcode histogram of every kernel in libavfe.so (no GPU needed): the proof of which Blackwell
instructions a kernel is made of -- UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (tensor-map TMA),
UBLKCP (1-D bulk TMA), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2 (fp32x2), IDP (dp2a), LDGSTS (cp.async).

    python profiles/sass_hist.py [out_dir] [kernel-substring ...]

Writes one `sass_hist_<kernel>.txt` per matching kernel (default: the hot kernels) and a one-line
summary per kernel to stdout.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "avsl_b200", "lib", "libavfe.so")
HOT = ["lip_frame_kernel", "lip_fused_kernel", "logmel_tile_kernel", "logmel_finalize_tiles_kernel", "pep_gemm2_kernel",
       "pep_gemm_kernel", "pep_stats_kernel", "fuse_vec_kernel", "fuse_ln_kernel", "fuse_ln_bwd_kernel", "logfbank_kernel",
       "noise_leaf_kernel", "noise_mix_kernel", "noise_cluster_kernel", "fuse_ln_tma_kernel", "fuse_ln_tma_tile_kernel", "gray_vec_kernel", "vfeats_kernel", "spec_time_warp_kernel"]
MARK = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "IDP", "LDGSTS", "DFMA", "DMUL", "DADD", "HMMA"]


def kernels():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, table = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            table[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", ln)
        if m and cur:
            table[cur][m.group(1)] += 1
    return table


def demangle(name):
    return subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name


def main():
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02")
    subs = sys.argv[2:] or HOT
    os.makedirs(out_dir, exist_ok=True)
    for name, ops in kernels().items():
        dm = demangle(name)
        hit = [s for s in subs if s in dm]
        if not hit:
            continue
        base = collections.Counter()
        for op, n in ops.items():
            base[op.split(".")[0]] += n
        tot = sum(ops.values())
        head = dm[:dm.index("(")] if "(" in dm and not dm.startswith("void avfe::lip") else dm.split("(")[0]
        if "<" in dm.split("(")[0]:
            head = dm[:dm.index(">(") + 1] if ">(" in dm else head
        short = re.sub(r"[^A-Za-z0-9_]+", "_", head.replace("void ", "").replace("avfe::", ""))[:90].strip("_")
        path = os.path.join(out_dir, f"sass_hist_{short}.txt")
        with open(path, "w") as f:
            f.write(f"# {dm}\n# {tot} SASS instructions (static count, cuobjdump -sass of avsl_b200/lib/libavfe.so)\n")
            f.write("# markers: " + ", ".join(f"{k}={base[k]}" for k in MARK if base[k]) + "\n")
            for op, n in ops.most_common():
                f.write(f"{n:6d}  {op}\n")
        print(f"{short:70s} {tot:6d} instrs  " + " ".join(f"{k}={base[k]}" for k in MARK if base[k]))


if __name__ == "__main__":
    main()
