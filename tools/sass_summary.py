"""profiles/sass_summary.txt: per-kernel SASS instruction histogram of libssdbox.so (cuobjdump -sass), the
evidence that the hot kernels are sm_100a code using TMA bulk copies (UBLKCP), mbarriers (SYNCS), cluster
barriers (UCGABAR), warp reductions (REDUX / MATCH / VOTE) and no tensor-core instructions (nothing on this path
is a contraction).

    python tools/sass_summary.py [path/to/libssdbox.so] > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "object-detection-pytorch_b200", "ssdbox", "lib", "libssdbox.so")
KEY = ["UBLKCP", "SYNCS", "UCGABAR", "REDUX", "MATCH", "VOTE", "SHFL", "ATOMS", "ATOMG", "RED", "MUFU", "LDS", "STS", "LDG",
       "LDGSTS", "STG", "BAR", "FFMA", "FFMA2", "FADD", "FADD2", "FMUL", "FMNMX", "FMNMX3", "DADD", "DFMA", "UTMALDG", "UTCHMMA", "UTCQMMA", "HMMA", "LDTM", "QGMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            if m.group(1) == "UBLKCP":
                cur["UBLKCP" + m.group(2)] += 1
    print("# SASS summary of %s" % os.path.relpath(LIB, ROOT))
    print("# architectures in the fatbin: %s" % ", ".join(arch))
    print("# per kernel: total instructions, then the count of each mnemonic of interest (absent = 0)")
    print("# UBLKCP = cp.async.bulk (1-D TMA; .S.G = global->shared load, .G.S = shared->global store), SYNCS = mbarrier,")
    print("# UCGABAR = barrier.cluster, REDUX / MATCH / VOTE = warp reductions / match.any / ballots.")
    print("# UTMALDG = cp.async.bulk.tensor (tensor-map TMA: the head-layout kernel's {32 positions x channels} boxes; every other")
    print("# tile on the path is a contiguous 1-D run), LDGSTS = cp.async; no UTC*MMA / HMMA / LDTM (no contraction on this path).")
    print()
    tot = collections.Counter()
    for name, c in kernels.items():
        n = sum(v for k, v in c.items() if "." not in k)
        items = ["%s=%d" % (k, c[k]) for k in KEY if c.get(k)]
        items += ["%s=%d" % (k, v) for k, v in sorted(c.items()) if k.startswith("UBLKCP.")]
        print("%-90s %6d  %s" % (name[:90], n, " ".join(items)))
        tot.update(c)
    print()
    print("TOTAL: " + " ".join("%s=%d" % (k, tot[k]) for k in KEY if tot.get(k)))
    print("tensor-core / tensor-map instructions: " + (" ".join("%s=%d" % (k, tot[k]) for k in ("UTMALDG", "UTCHMMA", "UTCQMMA", "HMMA", "LDTM", "QGMMA") if tot.get(k)) or "none"))


if __name__ == "__main__":
    main()
