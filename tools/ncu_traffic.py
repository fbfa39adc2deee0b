"""Turn the ncu metric pass over one step's GEMM launches into profiles/r02_gemm_traffic.json (read by bench.py for
`roofline.traffic`).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/r02_launches.csv python bench.py --kernel-only --no-graph --no-loss-check --steps 2 --warmup 3
    python tools/ncu_traffic.py gpurun_out/r02_gemm_traffic.csv profiles/r02_gemm_traffic.json
"""
import csv
import json
import sys
from collections import OrderedDict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
        "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    per = OrderedDict()
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        key = int(r[ci["ID"]])
        val = float(r[ci["Metric Value"]].replace(",", "")) * UNIT.get(r[ci["Metric Unit"]], 1.0)
        per.setdefault(key, {"name": r[ci["Kernel Name"]]})[r[ci["Metric Name"]]] = val
    # the launch list may cover every kernel of several steps: keep the GEMM launches of the LAST step (between the last
    # two adam_kernel launches), one-CTA (gemm_bf16_kernel) and two-CTA (gemm2_bf16_kernel) alike
    ids = sorted(per)
    adam = [i for i in ids if "adam_kernel" in per[i]["name"]]
    if len(adam) >= 2:
        ids = [i for i in ids if adam[-2] < i <= adam[-1]]
    gemm = [per[i] for i in ids if "gemm_bf16_kernel" in per[i]["name"] or "gemm2_bf16_kernel" in per[i]["name"]]
    n = len(gemm)
    rd = sum(v.get("dram__bytes_read.sum", 0.0) for v in gemm)
    wr = sum(v.get("dram__bytes_write.sum", 0.0) for v in gemm)
    t = sum(v.get("gpu__time_duration.sum", 0.0) for v in gemm)
    pair = sum(1 for v in gemm if "gemm2_bf16_kernel" in v["name"])
    out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the {n} GEMM launches ({pair} of them the "
                     f"two-CTA kernel) of one bench.py step (single_pass, B = 256); {src.split('/')[-1]}",
           "launches": n, "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
           "dram_bytes_per_launch": (rd + wr) / max(n, 1), "gpu_time_s_under_ncu": t}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
