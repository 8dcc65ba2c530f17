"""Summarise an `ncu --page source --csv` dump of conv3x3_tc_kernel: warp-stall samples per warp
role (prologue / TMA producer / MMA issuer / epilogue) and the hottest SASS instructions.
Usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_regions.py src.csv [block]"""
import csv
import sys


def blocks(path):
    out, cur = [], None
    for r in csv.reader(open(path)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            out.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    return out


def main():
    bl = blocks(sys.argv[1])
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    print(len(bl), "kernel blocks; using", k)
    b = bl[k]
    h = b["hdr"]
    ia, isamp, iexec, iaddr = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Address")
    stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    seen, seq = set(), []
    for r in b["rows"]:
        if r[iaddr] in seen:
            continue
        seen.add(r[iaddr])
        seq.append(r)
    tot = sum(int(r[isamp]) for r in seq)
    print(len(seq), "instructions,", tot, "samples")
    idx = lambda key: [i for i, r in enumerate(seq) if key in r[ia]]
    bars = idx("BAR.SYNC")
    reg = {"producer": idx("UTMALDG"), "mma": idx("UTCHMMA"), "epilogue": idx("LDTM")}
    order = sorted((v[0], v[-1], k) for k, v in reg.items() if v)
    bounds = [bars[0] + 1] + [max(o[0] - 60, bars[0] + 1) for o in order[1:]] + [bars[-1] - 3]
    spans = [("prologue", 0, bars[0] + 1)] + [(o[2], a, c) for o, a, c in zip(order, bounds[:-1], bounds[1:])] + \
            [("teardown", bars[-1] - 3, len(seq))]
    for name, a, c in spans:
        s, agg = 0, {}
        for r in seq[a:c]:
            s += int(r[isamp])
            for i in stall:
                if r[i] not in ("", "0"):
                    agg[h[i][6:]] = agg.get(h[i][6:], 0) + int(r[i])
        top = sorted(agg.items(), key=lambda x: -x[1])[:5]
        print(f"{name:9s} rows {a:5d}-{c:5d} samples {s:6d} ({100.0 * s / max(tot, 1):5.1f}%)  {top}")
    print("--- hottest instructions")
    for i, r in sorted(enumerate(seq), key=lambda x: -int(x[1][isamp]))[:22]:
        st = {h[j][6:]: int(r[j]) for j in stall if r[j] not in ("", "0")}
        print(f"{i:5d} {r[isamp]:>6s} exec {r[iexec]:>8s}  {r[ia].strip()[:72]:72s} {sorted(st.items(), key=lambda x: -x[1])[:2]}")


if __name__ == "__main__":
    main()
