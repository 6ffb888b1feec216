"""Write profiles/rollout_traffic.json from an `ncu --set full` capture of the rollout kernel (one launch) and tie it to the
CURRENT source of the kernel (bench.py reports `roofline.traffic` only while the hashes agree).

  python tools/update_traffic.py gpurun_out/prof_rollout_X.ncu-rep profiles/rX_rollout_ncu.md [batch] [nodes]
"""
import csv, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

rep, src_md = sys.argv[1], sys.argv[2]
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
nodes = int(sys.argv[4]) if len(sys.argv) > 4 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1}


def get(name):
    i = hdr.index(name)
    return float(vals[i].replace(",", "")) * scale.get(units[i], 1)


rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
j = {"kernel": vals[hdr.index("Kernel Name")].split("(")[0], "source": f"{src_md} (ncu --set full, one launch, EM-EVRP-{nodes}, batch {batch})",
     "source_sha256": bench.rollout_source_sha256(), "batch": batch, "nodes": nodes,
     "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
     "gpu__dram_throughput_pct_of_peak": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     "kernel_ms_under_ncu": get("gpu__time_duration.sum")}
with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "rollout_traffic.json"), "w") as f:
    json.dump(j, f, indent=1)
print(json.dumps(j, indent=1))
