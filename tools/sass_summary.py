"""Per-kernel SASS instruction census of the shipped extension (profiles/rNN_sass_summary.txt).
    python tools/sass_summary.py [out.txt]
Counts the mnemonics that show which hardware path a kernel uses: DMMA (fp64 tensor), HMMA (legacy mma.sync, here
TF32 m16n8k8 of the k-NN prefilter), LDGSTS (cp.async), LDS.128 / LDS.64 (fragment loads), UTMALDG (TMA), UTCxMMA /
LDTM (tcgen05 / TMEM), DFMA / DADD / DMUL (fp64 ALU), MUFU, BAR."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "cgp_b200", "_cgp_b200.so")
KEYS = ["DMMA", "HMMA", "LDGSTS", "LDS.128", "LDS.64", "UTMALDG", "UTC", "LDTM", "DFMA", "DADD", "DMUL", "MUFU", "BAR", "STG", "total"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", out)
    if names:
        dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
        demangle = dict(zip(names, dm))
    rows = []
    cur, cnt = None, None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                rows.append((cur, cnt))
            cur, cnt = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            cnt["total"] += 1
            for k in KEYS[:-1]:
                if op.startswith(k):
                    cnt[k] += 1
    if cur:
        rows.append((cur, cnt))

    def short(n):
        d = demangle.get(n, n)
        d = d.replace("void ", "")
        i = d.find(">(")
        d = d[:i + 1] if i >= 0 else re.sub(r"\(.*", "", d)
        return d.replace("(int)", "")

    agg = collections.OrderedDict()
    for n, c in rows:
        agg.setdefault(short(n), collections.Counter()).update(c)
    lines = ["# cuobjdump -sass cgp_b200/_cgp_b200.so (sm_100a), instruction counts per kernel (template instances listed separately)",
             "# DMMA = fp64 tensor core (mma.sync.m8n8k4.f64); HMMA = mma.sync TF32 (k-NN prefilter); LDGSTS = cp.async; "
             "UTMALDG = TMA; UTC*/LDTM = tcgen05/TMEM (fp64 has no tcgen05 kind)",
             f"{'kernel':58s} " + " ".join(f"{k:>8s}" for k in KEYS)]
    tot = collections.Counter()
    for n, c in sorted(agg.items()):
        lines.append(f"{n[:58]:58s} " + " ".join(f"{c[k]:8d}" for k in KEYS))
        tot.update(c)
    lines.append(f"{'ALL KERNELS':58s} " + " ".join(f"{tot[k]:8d}" for k in KEYS))
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
