"""Summarises one `ncu --set full` capture of the bench workload into profiles/ncu_traffic.json: per kernel family the
DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of its LARGEST launch, which bench.py reports as
`roofline.traffic`.   python tools/ncu_traffic.py gpurun_out/<capture>.ncu-rep [git-rev]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = {"leaf": "leaf_tree_kernel", "bc_round0": "r0", "ntt_pass": "ntt_strided_pass_kernel", "ntt_final": "ntt_final_pass_kernel"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

rep = sys.argv[1]
rev = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
res = {}
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    name = d.get("Kernel Name", "")
    for fam, pat in FAMILIES.items():
        if pat in name:
            rd = float(d["dram__bytes_read.sum"]) * UNIT[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * UNIT[u["dram__bytes_write.sum"]]
            ms = float(d["gpu__time_duration.sum"]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u["gpu__time_duration.sum"], 1)
            if fam not in res or rd + wr > res[fam]["bytes_per_launch"]:
                res[fam] = {"bytes_per_launch": int(rd + wr), "read": int(rd), "write": int(wr), "ms_under_ncu": ms, "grid": d.get("Grid Size"),
                            "kernel": name.split("(")[0],
                            "note": f"largest launch of the family in {os.path.basename(rep)} (ncu --set full --clock-control none, commit {rev}): "
                                    f"{rd / 1e9:.3f} GB read + {wr / 1e9:.3f} GB written"}
json.dump(res, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
