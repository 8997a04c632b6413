#!/usr/bin/env python3
"""dram bytes per launch of every kernel from an `ncu --set full` raw-page CSV -> profiles/rNN_traffic.json
(the file bench.py reads for `roofline.traffic`).

  ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
  python tools/ncu_traffic.py raw.csv "note" > profiles/rNN_traffic.json
"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, note):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in data:
        name = r[ix["Kernel Name"]].replace("<unnamed>::", "").replace("void ", "").split("(")[0].replace("(int)", "")
        if name.startswith("at::"):
            continue
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[ix[m]].replace(",", "")) * UNIT.get(units[ix[m]], 1.0)
        out.setdefault(name, []).append({"grid": int(float(r[ix["launch__grid_size"]].replace(",", ""))), "dram_bytes": b,
                                         "dur_us": float(r[ix["gpu__time_duration.sum"]].replace(",", "")) *
                                         {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1.0)})
    # one entry per distinct (kernel, grid)
    for k in out:
        seen, uniq = set(), []
        for e in out[k]:
            if e["grid"] not in seen:
                seen.add(e["grid"]); uniq.append(e)
        out[k] = uniq
    print(json.dumps({"note": note, "kernels": out}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
