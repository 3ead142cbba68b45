"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals after engine.prepare()."""
import collections
import csv
import re
import sys


def main(path, out=None):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        m = (re.search(r"<unnamed>::(\w+)", name) or re.search(r"\(anonymous namespace\)::(\w+)", name)
             or re.search(r"\brp2::(\w+)", name))
        rows.append((m.group(1) if m and not name.startswith("void at::") and "at::" not in name[:60] else "torch/other", v))
    idx = max((i for i, (n, _) in enumerate(rows) if n == "sumsq_kernel"), default=0)
    body = rows[idx:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in body:
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"{path}: {len(body)} launches after prepare(), total {tot / 1e3:.2f} ms (cold-cache, serialised: compare SHARES)"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"  {t / 1e3:9.3f} ms {100 * t / tot:5.1f}%  x{c:4d}  avg {t / c:9.1f} us  {n}")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
