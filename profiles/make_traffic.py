"""profiles/traffic.json from an `ncu --page raw --csv` dump of the bench kernels:
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, keyed by bench.py's kernel name."""
import csv
import json
import re
import sys

KEYS = {("1", "2", "1"): "q1_stiffness", ("0", "2", "1"): "q1_mass", ("1", "2", "2"): "q2_stiffness", ("0", "2", "2"): "q2_mass"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, out):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    res = {}
    for r in data:
        m = re.search(r"FormTraits<(\d+), (\d+), (\d+)", r[ik])
        if not m or m.groups() not in KEYS:
            continue
        res[KEYS[m.groups()]] = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
    res["_source"] = "ncu --set full --clock-control none (one launch each); " + path.split("/")[-1]
    json.dump(res, open(out, "w"), indent=1)
    print(res)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
