#!/usr/bin/env python
"""DRAM traffic per row of the hot kernels, from an `ncu --set full` capture -> profiles/traffic.json (read by bench.py for
`roofline.traffic`).

    python tools/ncu_traffic.py REPORT.ncu-rep CONFIG ROWS [KERNEL_REGEX=timer ...]

e.g.  python tools/ncu_traffic.py gpurun_out/r2g_prof.ncu-rep M 10000000 knn_tile=knn_search features_direct=features
The file maps config -> timer name -> (dram__bytes_read.sum + dram__bytes_write.sum) / rows of ONE launch, plus the source."""
import csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, config, rows = sys.argv[1], sys.argv[2], int(sys.argv[3])
    maps = [a.split("=") for a in sys.argv[4:]]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    table = list(csv.reader(out.splitlines()))
    hdr, units = table[0], table[1]
    col = {h: i for i, h in enumerate(hdr)}
    path = os.path.join(ROOT, "profiles", "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    entry = data.setdefault(config, {})
    for r in table[2:]:
        name = r[col["Kernel Name"]]
        for rx, timer in maps:
            if re.search(rx, name) and timer not in entry.get("_seen", []):
                tot = 0.0
                for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(r[col[m]].replace(",", "")) * UNIT[units[col[m]]]
                entry[timer] = tot / rows
                entry.setdefault("_seen", []).append(timer)
    entry.pop("_seen", None)
    entry["_source"] = os.path.basename(rep) + " (ncu --set full, one launch per kernel, %d rows)" % rows
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data[config], indent=1))


if __name__ == "__main__":
    main()
