#!/usr/bin/env python
"""Dynamic instruction counts, stall samples and shared-memory wavefronts per KERNEL source line.

Joins `nvdisasm -gi <cubin>` (static SASS with //## File ... line N [inlined at ...] chains; the last marker
of a chain is the line of the kernel body itself) with
`ncu -i X.ncu-rep --page source --csv --print-source sass` by instruction order.

    python profiles/line_profile.py all_gi.dis sass.csv <mangled kernel name> [bucket]
bucket = number of kernel source lines merged per output row (default 1).
"""
import collections, csv, re, sys


def static_lines(dis, func):
    out, on, chain = [], False, []
    for ln in open(dis):
        if ln.startswith(".text."):
            on = (ln.strip() == f".text.{func}:")
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            chain.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m:
            toks = m.group(2).split()
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            if chain:
                cur = chain[-1]
            out.append((cur, op.split(".")[0]))
            chain = []
    return out


def dynamic(path):
    rows, hdr, seen = [], None, 0
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            seen += 1
            if seen > 1:
                break
            hdr = None
            continue
        if hdr is None:
            hdr = row
            continue
        d = dict(zip(hdr, row))
        toks = d["Source"].split()
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        st = {k[6:]: int(v or 0) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k}
        rows.append((op.split(".")[0], int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0),
                     int(d.get("L1 Wavefronts Shared") or 0), st))
    return rows


def main():
    dis, sass, func = sys.argv[1:4]
    bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    st, dy = static_lines(dis, func), dynamic(sass)
    assert len(st) == len(dy), (len(st), len(dy))
    agg = collections.OrderedDict()
    tot = sum(d[1] for d in dy); stot = sum(d[2] for d in dy) or 1; wtot = sum(d[3] for d in dy) or 1
    for (loc, op), (_, n, s, w, reasons) in zip(st, dy):
        key = (loc[0], loc[1] // bucket * bucket)
        a = agg.setdefault(key, [0, 0, 0, 0, collections.Counter()])
        a[0] += 1; a[1] += n; a[2] += s; a[3] += w; a[4].update(reasons)
    print(f"# {len(st)} instructions, executed {tot}, samples {stot}, smem wavefronts {wtot}")
    print(f"{'file':18s} {'line':>5s} {'static':>6s} {'exec%':>6s} {'smp%':>6s} {'smemwf%':>7s}  top stall reasons (share of this row's samples)")
    for key in sorted(agg, key=lambda k: (k[0], k[1])):
        a = agg[key]
        if a[1] * 500 >= tot or a[2] * 500 >= stot or a[3] * 200 >= wtot:
            rs = ", ".join(f"{k} {v / max(sum(a[4].values()), 1) * 100:.0f}%" for k, v in a[4].most_common(4))
            print(f"{key[0]:18s} {key[1]:5d} {a[0]:6d} {a[1] / tot * 100:6.1f} {a[2] / stot * 100:6.1f} {a[3] / wtot * 100:7.1f}  {rs}")


if __name__ == "__main__":
    main()
