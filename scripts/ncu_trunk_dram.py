#!/usr/bin/env python
"""dram__bytes_read.sum + dram__bytes_write.sum per launch of the tower kernel from an `ncu --set full` report ->
profiles/r02_trunk4_dram.json, which bench.py reads for roofline.traffic (no pasted literals).

    python scripts/ncu_trunk_dram.py gpurun_out/prof_trunk_r02.ncu-rep [kernel-regex]
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else "k_trunk4")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    per = []
    for r in rows[2:]:
        if not pat.search(r[kn]):
            continue
        rd = float(r[ir].replace(",", "")) * UNIT[units[ir]]
        wr = float(r[iw].replace(",", "")) * UNIT[units[iw]]
        per.append({"read": rd, "write": wr, "duration": r[it] + " " + units[it], "grid": r[hdr.index("launch__grid_size")]})
    if not per:
        sys.exit("no launch matching %s in %s" % (pat.pattern, rep))
    mean = sum(p["read"] + p["write"] for p in per) / len(per)
    doc = {"dram_bytes_per_launch": mean, "read_bytes_per_launch": sum(p["read"] for p in per) / len(per),
           "write_bytes_per_launch": sum(p["write"] for p in per) / len(per), "launches": per,
           "source": "ncu --set full --clock-control none, %s, mean of %d captured launches (%s); cold-cache, serialised replays"
                     % (os.path.basename(rep), len(per), pat.pattern)}
    with open(os.path.join(ROOT, "profiles", "r02_trunk4_dram.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc)[:400])


if __name__ == "__main__":
    main()
