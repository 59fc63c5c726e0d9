#!/usr/bin/env python
"""profiles/ncu_traffic_r1.json from the raw pages of the `ncu --set full` captures that
scripts/gpu_profile_all.sh leaves in gpurun_out/ (dram__bytes_read.sum + dram__bytes_write.sum per launch).
Usage: scripts/make_traffic.py TAG family=kernel_regex_name ...   e.g. r1h loo_em=loo_em_step5"""
import csv
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def main():
    tag = sys.argv[1]
    out = {"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures at 200k sites x 500 x 10 "
                   "(scripts/gpu_profile_all.sh %s); bench.py scales them linearly with the site count." % tag, "kernels": {}}
    for arg in sys.argv[2:]:
        fam, kern = arg.split("=")
        path = os.path.join(ROOT, "gpurun_out", "raw_%s_%s.csv" % (kern, tag))
        rows = list(csv.reader(io.StringIO(open(path).read())))
        hdr, units, r = rows[0], rows[1], rows[2]

        def val(key):
            i = hdr.index(key)
            return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        out["kernels"][fam] = {"kernel": r[hdr.index("Kernel Name")][:60], "sites": 200000, "individuals": 500, "populations": 10,
                               "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                               "ncu_time_s": val("gpu__time_duration.sum"), "source": os.path.basename(path)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic_%s.json" % tag.split("_")[0]), "w"), indent=1)
    print(json.dumps(out["kernels"], indent=1))


if __name__ == "__main__":
    main()
