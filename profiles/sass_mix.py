#!/usr/bin/env python
"""Dynamic SASS opcode mix per kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`."""
import collections, csv, sys
def main(path, top=28):
    kernels, cur = [], None
    for row in csv.reader(open(path)):
        if not row: continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}; kernels.append(cur); continue
        if cur is None: continue
        if cur["hdr"] is None: cur["hdr"] = row; continue
        cur["rows"].append(row)
    seen = set()
    for k in kernels:
        if k["name"] in seen:           # several launches of one kernel: the first is enough
            continue
        seen.add(k["name"])
        h = k["hdr"]; iS, iE, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        mix, smp = collections.Counter(), collections.Counter(); tot = 0
        for r in k["rows"]:
            toks = r[iS].split()
            if not toks: continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            op = op.split(".")[0]
            n = int(r[iE] or 0); mix[op] += n; tot += n; smp[op] += int(r[iSm] or 0)
        stot = sum(smp.values()) or 1
        print(f"## {k['name'][:70]}: {tot} warp-instructions, {len(k['rows'])} static")
        for op, n in mix.most_common(top):
            print(f"  {op:12s} {n:12d} {n / tot * 100:5.1f}%   samples {smp[op] / stot * 100:5.1f}%")
        # opcodes that prove which hardware paths ran, whatever their rank: TMA tile loads, tensor-memory loads / stores,
        # tcgen05.mma, warp reductions, packed FP32, local-memory (spill) traffic
        static = collections.Counter()
        for r in k["rows"]:
            toks = r[iS].split()
            if toks:
                op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
                static[op.split(".")[0]] += 1
        census = ["UTMALDG", "UBLKCP", "SYNCS", "LDTM", "STTM", "UTCHMMA", "UTCBAR", "CREDUX", "REDUX", "FADD2", "FMUL2", "FFMA2",
                  "I2F", "MUFU", "DFMA", "LDL", "STL", "BAR"]
        print("  census (static / executed): " + ", ".join(f"{op} {static[op]}/{mix[op]}" for op in census if static[op]))
if __name__ == "__main__":
    main(sys.argv[1])
