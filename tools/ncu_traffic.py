"""dram bytes per stage-kernel launch from an ncu --set full report that holds the 4 stage launches of one step:
python tools/ncu_traffic.py report.ncu-rep profiles/stage_kernel_traffic.json"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    i = hdr.index(name)
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[units[i]]
    return [float(r[i].replace(",", "")) * scale for r in data]
rd, wr, t = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
names = [r[hdr.index("Kernel Name")] for r in data]
per = [a + b for a, b in zip(rd, wr)]
res = {"source": rep.split("/")[-1], "launches": len(per), "kernels": names,
       "dram_bytes_read": rd, "dram_bytes_write": wr, "duration_s_under_ncu": t,
       "dram_bytes_per_launch": sum(per) / len(per)}
json.dump(res, open(out, "w"), indent=1)
print(res["dram_bytes_per_launch"], [round(x / 1e9, 3) for x in per])
