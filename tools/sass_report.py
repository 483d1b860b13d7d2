"""Static evidence from the built library (no GPU needed): per kernel of csrc/libvqvae_b200.so the counts of the SASS opcodes
that show HOW it talks to the hardware — UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tensor-memory loads / stores), UTCBAR
(tcgen05.commit), UTMALDG / UTMASTG (TMA tensor loads / stores), UBLKCP (bulk copies), UTMAPF (TMA L2 prefetch), SYNCS
(mbarrier ops) — plus registers / spills / shared memory from the ptxas logs.  Writes profiles/<round>_sass_histogram.txt and
profiles/<round>_ptxas.txt.   Usage: python tools/sass_report.py r2"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vae-based-music--deep-generative-models_b200", "csrc")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "UBLKPF", "SYNCS", "LDGSTS", "REDUX", "CREDUX"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main(rnd):
    so = os.path.join(CSRC, "libvqvae_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            per[cur]["_total"] += 1
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    per[cur][o] += 1
    dm = demangle(list(per))
    lines = [f"SASS opcode counts per kernel of libvqvae_b200.so (cuobjdump -sass, sm_100a) — {rnd}", "",
             f"{'kernel':78s} {'instrs':>7s} " + " ".join(f"{o:>8s}" for o in OPS)]
    tot = collections.Counter()
    for k, c in per.items():
        name = re.sub(r"^void ", "", dm[k])
        name = re.sub(r"\(.*", "", name)[:78]
        lines.append(f"{name:78s} {c['_total']:7d} " + " ".join(f"{c[o]:8d}" for o in OPS))
        tot.update(c)
    lines.append(f"{'TOTAL':78s} {tot['_total']:7d} " + " ".join(f"{tot[o]:8d}" for o in OPS))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"{rnd}_sass_histogram.txt"), "w").write("\n".join(lines) + "\n")
    # ptxas: registers / spills / smem per entry point
    out = [f"ptxas -v summary per kernel (nvcc -O3 -gencode arch=compute_100a,code=sm_100a) — {rnd}", ""]
    for log in sorted(glob.glob(os.path.join(CSRC, "*.ptxas.log"))):
        txt = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                             r"ptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt):
            name = re.sub(r"\(.*", "", re.sub(r"^void ", "", demangle([m.group(1)])[m.group(1)]))[:84]
            out.append(f"{os.path.basename(log)[:-10]:14s} {name:84s} regs {int(m.group(5)):3d}  spill st/ld {int(m.group(3)):4d}/{int(m.group(4)):4d} B  stack {int(m.group(2)):4d} B")
    open(os.path.join(ROOT, "profiles", f"{rnd}_ptxas.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(lines[-1:]))
    print("kernels:", len(per), "ptxas entries:", len(out) - 2)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2")
