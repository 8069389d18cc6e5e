"""Record the measured DRAM traffic of a GEMM signature for bench.py's `roofline.traffic`:
   python tools/ncu_traffic.py <capture.ncu-rep> <epilogue> <M> <N> <K> [source note]
reads dram__bytes_read.sum + dram__bytes_write.sum of the first kernel in an `ncu --set full` capture and stores it under
"<epilogue>,<N>,<K>" in profiles/ncu_traffic.json (bench.py scales it linearly in M)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, epi, m, n, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    note = sys.argv[6] if len(sys.argv) > 6 else os.path.basename(rep)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    total = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        total += float(vals[i].replace(",", "")) * UNIT[units[i]]
    kernel = vals[hdr.index("Kernel Name")]
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tbl = json.load(open(path)) if os.path.exists(path) else {}
    tbl["%d,%d,%d" % (epi, n, k)] = {"m": m, "bytes": total, "source": "%s: %s" % (note, kernel[:80])}
    json.dump(tbl, open(path, "w"), indent=1, sort_keys=True)
    print(path, tbl["%d,%d,%d" % (epi, n, k)])


if __name__ == "__main__":
    main()
