"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list into a markdown table (one forward pass)."""
import collections, csv, re, sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else "launch list"
lines = [l for l in open(path) if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("unnamed>::", "").replace("<unnamed>::", "")
    seq.append((name, v, row["Grid Size"], row["Block Size"]))
starts = [i for i, s in enumerate(seq) if "stft_power" in s[0]]
one = seq[starts[-2]:starts[-1]] if len(starts) >= 2 else seq
tot = sum(s[1] for s in one)
agg = collections.OrderedDict()
for n, v, g, b in one:
    a = agg.setdefault(n, [0, 0.0, g, b])
    a[0] += 1
    a[1] += v
print(f"# {title}\n")
print(f"One forward pass (32 x 10 s): {len(one)} launches, {tot:.1f} us summed under ncu (cold-cache, serialised: read the SHARES).\n")
print("| kernel | launches | total us | share | avg us | grid | block |\n|---|---|---|---|---|---|---|")
for n, (c, v, g, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{n}` | {c} | {v:.1f} | {100 * v / tot:.1f}% | {v / c:.1f} | {g} | {b} |")
