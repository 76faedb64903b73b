"""profiles/traffic.json from an `ncu --set full` capture of the headline kernel at config 2:
    python tools/make_traffic.py gpurun_out/<capture>.ncu-rep profiles/<summary>.csv
The hash of the kernel sources goes in with it: bench.py reports `roofline.traffic` only when it runs those sources."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ncu_summary as N

import bench

rep, out_csv = sys.argv[1], sys.argv[2]
rows = N.ncu_csv(rep, "raw")
hdr, units, r = rows[0], rows[1], rows[2]


def val(name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
N.summary(rep, out_csv)
d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
     "algorithmic_bytes_per_launch": bench.BYTES_PER_FRAME_FUSED_ADD * bench.SCENES * bench.FRAMES, "source": out_csv + " (ncu --set full, config 2 launch)",
     "csrc_hash": bench.csrc_hash()}
with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w") as f:
    json.dump(d, f, indent=1)
print(d)
