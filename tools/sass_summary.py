"""Counts the Blackwell-only SASS mnemonics per kernel of csrc/libdrl_b200.so (cuobjdump -sass): UTCHMMA (tcgen05.mma),
LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UTMALDG (cp.async.bulk.tensor), UBLKCP (cp.async.bulk), SYNCS (mbarrier),
UTCATOMSWS / UTCALLOC-style TMEM allocation.   python tools/sass_summary.py > profiles/r02_sass_tcgen05.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "distributed_reinforcement_learning_b200", "csrc", "libdrl_b200.so")
PAT = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "SYNCS", "ELECT", "UTCATOMSWS", "R2UR"]


def demangle(names):
    p = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return p.stdout.split("\n")


def short(n):
    n = re.sub(r"\bdrl::", "", n)
    n = re.sub(r"umma16::", "", n)
    n = re.sub(r"\(int\)|\(bool\)", "", n)
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(T\d.*$", "", n)
    return n[:150]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for p in PAT:
            if re.search(r"\b" + p + r"[\.\s]", line):
                counts[cur][p] += 1
    names = list(counts)
    dem = demangle(names)
    rows = [(short(d), counts[n]) for n, d in zip(names, dem) if counts[n]["UTCHMMA"] or counts[n]["UTMALDG"] or counts[n]["UBLKCP"]]
    print("# cuobjdump -sass csrc/libdrl_b200.so: Blackwell-only mnemonics per kernel (static instruction counts)")
    print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA tensor load),")
    print("# UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, ELECT = elect.sync.  %d kernels with tensor-core or bulk-copy code.\n" % len(rows))
    print("| kernel | " + " | ".join(PAT) + " |")
    print("|---|" + "---|" * len(PAT))
    tot = collections.Counter()
    for nm, c in sorted(rows):
        tot.update(c)
        print("| `%s` | %s |" % (nm, " | ".join(str(c[p]) for p in PAT)))
    print("| **total** | %s |" % " | ".join(str(tot[p]) for p in PAT))


if __name__ == "__main__":
    main()
