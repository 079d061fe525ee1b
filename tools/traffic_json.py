"""DRAM bytes and duration per launch from an .ncu-rep -> profiles/<tag>_traffic.json (read by bench.py).
Usage: traffic_json.py <rep> <out.json> [note] [plan-info.json written by profile_step.py --info-out]"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def val(r, key):
    v = float(r[ix[key]].replace(",", ""))
    u = units[ix[key]].lower()
    return v * {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
launches = [{"kernel": r[ix["Kernel Name"]][:60], "grid": r[ix["launch__grid_size"]],
             "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
             "duration_us": val(r, "gpu__time_duration.sum")} for r in data]
doc = {"source": "ncu --set full --clock-control none, tools/profile_step.py (1 M block in ground contact, 2nd frame)"
                 + (", " + sys.argv[3] if len(sys.argv) > 3 else ""), "launches": launches}
if len(sys.argv) > 4:
    plan = json.load(open(sys.argv[4]))
    doc.update(rounds_per_sweep=plan["rounds_per_sweep"], tiles_in_pass=plan["tiles_in_pass"], substeps=plan["substeps"], iterations=plan["iterations"])
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(launches[0]))
