"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel uses (tcgen05 = UTCHMMA / UTCQMMA,
tensor-memory loads / stores = LDTM / STTM, TMA = UTMALDG / UTMASTG, legacy mma.sync = HMMA), from the shipped library.

    python tools/sass_summary.py avsiam_b200/libavsiam_b200.so > profiles/r02_sass_summary.txt
"""
import re
import subprocess
import sys

MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "HMMA", "MUFU.EX2", "MUFU.TANH", "SYNCS"]


def main():
    lib = sys.argv[1]
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", "-p", n], capture_output=True, text=True).stdout.strip()
    cur, counts, order = None, {}, []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = {k: 0 for k in MNEMONICS}
            counts[cur]["instructions"] = 0
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        mm = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not mm:
            continue
        op = mm.group(1)
        counts[cur]["instructions"] += 1
        for k in MNEMONICS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
    print(f"# {lib}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
    print("# " + " ".join(f"{k:>9s}" for k in ["instr"] + MNEMONICS) + "  kernel")
    tot = {k: 0 for k in MNEMONICS}
    for fn in order:
        c = counts[fn]
        for k in MNEMONICS:
            tot[k] += c[k]
        name = demangle(fn).replace("(anonymous namespace)::", "").replace("void ", "")
        print("  " + " ".join(f"{c[k]:9d}" for k in ["instructions"] + MNEMONICS) + "  " + name[:110])
    print("# total " + " ".join(f"{k}={v}" for k, v in tot.items()))


if __name__ == "__main__":
    main()
