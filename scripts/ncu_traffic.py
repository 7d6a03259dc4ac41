"""profiles/r2_traffic.json from an `ncu --set full` raw CSV of ONE backbone forward (scripts/one_forward.py, chain launches
only): DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per kernel family and step.  The first 10 chain launches of a
forward are the SA family (SA1, SA2, SA3: one launch per scale; SA4: two per scale), the rest the FP family.
usage: python scripts/ncu_traffic.py gpurun_out/<raw>.csv [source note]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
I = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


fam = {"sa_mlp": 0.0, "fp_mlp": 0.0}
launches = []
for k, r in enumerate(data):
    b = sum(to_bytes(r[I[m]], units[I[m]]) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    f = "sa_mlp" if k < 10 else "fp_mlp"
    fam[f] += b
    launches.append({"kernel": r[I["Kernel Name"]][:60], "grid": r[I["launch__grid_size"]], "us": float(r[I["gpu__time_duration.sum"]]),
                     "dram_read_MB": to_bytes(r[I["dram__bytes_read.sum"]], units[I["dram__bytes_read.sum"]]) / 1e6,
                     "dram_write_MB": to_bytes(r[I["dram__bytes_write.sum"]], units[I["dram__bytes_write.sum"]]) / 1e6, "family": f})
out = {"sa_mlp": fam["sa_mlp"], "fp_mlp": fam["fp_mlp"], "unit": "bytes per step (batch of 16 scenes), summed over the family's launches",
       "_source": (sys.argv[2] if len(sys.argv) > 2 else os.path.basename(sys.argv[1])) + ": ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum",
       "launches": launches}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print({k: round(v / 1e6, 1) for k, v in fam.items()}, "MB;", len(launches), "launches")
