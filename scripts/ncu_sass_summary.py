"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` output: per kernel, instruction mix by opcode
(executed warp-instructions) and stall-sample totals by reason, plus the hottest instructions.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source sass > /tmp/sass.csv
    python scripts/ncu_sass_summary.py /tmp/sass.csv [kernel-substring] [top-N]
"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
only = int(sys.argv[4]) if len(sys.argv) > 4 else -1  # n-th matching launch only
seen = -1
rows = list(csv.reader(open(path)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(rows[j])
            j += 1
        i = j
        if want not in name or not body:
            continue
        seen += 1
        if only >= 0 and seen != only:
            continue
        ix = {h: k for k, h in enumerate(hdr)}
        ops = defaultdict(int)
        stalls = defaultdict(int)
        tot_inst = tot_samp = 0
        for r in body:
            src = r[ix["Source"]].strip()
            toks = src.split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            op = op.split(".")[0]
            n = int(r[ix["Instructions Executed"]] or 0)
            ops[op] += n
            tot_inst += n
            tot_samp += int(r[ix["# Samples"]] or 0)
            for h in hdr:
                if h.startswith("stall_") and "Not Issued" not in h:
                    stalls[h] += int(r[ix[h]] or 0)
        print("=" * 100)
        print(name[:110])
        print(f"warp-instructions {tot_inst}   samples {tot_samp}")
        print("opcode mix:", ", ".join(f"{k} {v / tot_inst:.1%}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:22]))
        print("stalls:", ", ".join(f"{k[6:]} {v / max(tot_samp, 1):.1%}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]))
        print("hottest instructions (samples, executed, top stall, sass):")
        for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:topn]:
            st = max(((h, int(r[ix[h]] or 0)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h), key=lambda kv: kv[1])
            print(f"  {int(r[ix['# Samples']]):7d} {int(r[ix['Instructions Executed']]):10d} {st[0][6:]:>12s}  {r[ix['Source']].strip()[:90]}")
    else:
        i += 1
