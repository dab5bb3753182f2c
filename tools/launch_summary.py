"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: launches, total ms, share.
    python tools/launch_summary.py profiles/r02_launches_bench.csv.gz > profiles/r02_launches_bench_summary.txt"""
import collections
import csv
import gzip
import io
import re
import sys


def main(path):
    raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    lines = [l for l in raw.splitlines() if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("\n".join(lines))))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = r[ik]
        m = re.match(r"(?:void )?([A-Za-z_0-9:]+)", name)
        short = m.group(1) if m else name[:40]
        if short.startswith("at::") or short.startswith("cutlass") or "cublas" in short.lower():
            short = short.split("<")[0][:28]
        v = float(r[iv].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        tot[short][0] += 1
        tot[short][1] += ms
    total = sum(v[1] for v in tot.values())
    n = sum(v[0] for v in tot.values())
    print(f"# total {total / 1e3:.3f} s over {n} launches")
    print(f"{'kernel':28s} {'launches':>9s} {'total_ms':>11s} {'share_%':>8s} {'avg_us':>10s}")
    for k, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:28s} {c:9d} {ms:11.2f} {100 * ms / total:8.2f} {1e3 * ms / c:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
