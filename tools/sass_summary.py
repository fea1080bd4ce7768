"""Static evidence that needs no GPU: per-kernel registers / spills (ptxas -v) and the count of tensor-core, TMEM, TMA and
mbarrier SASS instructions in each kernel of each .cu (cuobjdump -sass).  python tools/sass_summary.py > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "egocentric-gaze-prediction_b200", "csrc")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# mnemonic prefixes worth counting: UTCHMMA/UTCQMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st,
# UTCATOMSWS = TMEM alloc, UTMALDG/UTMASTG = TMA tensor load/store, SYNCS = mbarrier ops, UBLKCP = bulk copy
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS",
         "HMMA", "LDGSTS", "RED", "ATOM", "LDL", "STL")


def demangle(names):
    out = subprocess.run([os.path.join(CUDA, "bin", "cu++filt")] + names, capture_output=True, text=True).stdout.splitlines()
    out = [re.sub(r"^void ", "", n).replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("(int)", "") for n in out]
    return [re.sub(r"\(.*", "", n) for n in out]


def main():
    tmp = tempfile.mkdtemp()
    for src in sorted(f for f in os.listdir(CSRC) if f.endswith(".cu")):
        obj = os.path.join(tmp, src[:-3] + ".o")
        res = subprocess.run([os.path.join(CUDA, "bin", "nvcc")] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj],
                             capture_output=True, text=True)
        if res.returncode != 0:
            sys.exit(res.stderr)
        # ptxas -v: "Compiling entry function 'X'" / "N bytes stack frame, S bytes spill stores, L bytes spill loads" / "Used R registers"
        regs = {}
        cur = None
        for ln in res.stderr.splitlines():
            m = re.search(r"Compiling entry function '([^']+)'", ln)
            if m:
                cur = m.group(1)
                regs[cur] = {"regs": None, "spill": (0, 0), "smem": 0}
            m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m and cur:
                regs[cur]["spill"] = (int(m.group(1)), int(m.group(2)))
            m = re.search(r"Used (\d+) registers", ln)
            if m and cur:
                regs[cur]["regs"] = int(m.group(1))
        sass = subprocess.run([os.path.join(CUDA, "bin", "cuobjdump"), "-sass", obj], capture_output=True, text=True).stdout
        counts = collections.OrderedDict()
        cur = None
        for ln in sass.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                cur = m.group(1)
                counts[cur] = collections.Counter()
                continue
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", ln)
            if m and cur:
                op = m.group(1)
                counts[cur]["_total"] += 1
                for w in WATCH:
                    if op.startswith(w):
                        key = op if w in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTCBAR") else w
                        counts[cur][key] += 1
        names = list(counts)
        pretty = dict(zip(names, demangle(names))) if names else {}
        print("== %s" % src)
        for n in names:
            r = regs.get(n, {"regs": None, "spill": (0, 0)})
            c = counts[n]
            ops = ", ".join("%s x%d" % (k, v) for k, v in sorted(c.items()) if k != "_total")
            print("  %-58s regs %3s  spill st/ld %d/%d B  %5d SASS instr%s" % (
                pretty.get(n, n)[:58], r["regs"], r["spill"][0], r["spill"][1], c["_total"], ("  | " + ops) if ops else ""))


if __name__ == "__main__":
    main()
