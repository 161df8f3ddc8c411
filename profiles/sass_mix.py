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
    for k in kernels:
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
if __name__ == "__main__":
    main(sys.argv[1])
