"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py: the kernels of the LAST step
(between the last two adam_kernel launches), grouped by kernel, with their share of the step.

    python tools/launch_summary.py gpurun_out/r02_launches.csv > profiles/r02_launches_summary.txt
"""
import csv
import re
import sys
from collections import OrderedDict

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "s": 1e6, "second": 1e6}


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([\w:]+(?:<[^()]*?>)?)", name)
    return (m.group(1) if m else name)[:90]


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    launches = [(r[ci["Kernel Name"]], float(r[ci["Metric Value"]].replace(",", "")) * UNIT.get(r[ci["Metric Unit"]], 1.0))
                for r in rows[1:] if len(r) == len(hdr) and r[ci["Metric Name"]] == "gpu__time_duration.sum"]
    adam = [i for i, (n, _) in enumerate(launches) if "adam_kernel" in n]
    if len(adam) >= 2:
        step = launches[adam[-2] + 1:adam[-1] + 1]
    else:
        step = launches
    agg = OrderedDict()
    for n, us in step:
        a = agg.setdefault(short(n), [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"# {sys.argv[1]}: kernels of the last bench.py step ({len(step)} launches, {tot / 1e3:.2f} ms under ncu: cold caches,")
    print("# serialised launches — compare SHARES with bench.py's kernel_families_ms, not absolute times)")
    print(f"# {'launches':>8s} {'ms':>9s} {'share':>7s}  kernel")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {c:8d} {us / 1e3:9.3f} {100 * us / tot:6.2f}%  {k}")


if __name__ == "__main__":
    main()
