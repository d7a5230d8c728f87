"""profiles/bench_traffic.json from an ncu --csv metrics log of the steady-state launch of the bench workload
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... --launch-skip 3 -c 1 --csv --log-file X python
scripts/prof_level.py 0 exact 2`): bench.py reports it as roofline.traffic (DRAM bytes of ONE launch of the timed kernel)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1]
vals = {}
for row in csv.reader(l for l in open(src) if l.startswith('"')):
    if len(row) > 14 and row[0] != "ID":
        name, unit, v = row[12], row[13], float(row[14].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1,
                 "sector": 1}.get(unit, 1)
        vals[name] = v * scale
        kernel = row[4]
out = {"dram_bytes_per_launch": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
       "dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals["dram__bytes_write.sum"],
       "kernel_seconds_under_ncu": vals.get("gpu__time_duration.sum"), "kernel": kernel, "think_ms": 0.0,
       "l2_sectors_read": vals.get("lts__t_sectors_op_read.sum"), "l2_sectors_write": vals.get("lts__t_sectors_op_write.sum"),
       "l2_sectors_atom": vals.get("lts__t_sectors_op_atom.sum"),
       "source": f"profiles/{os.path.basename(src)}: ncu metrics of the steady-state launch (4th launch of the process: buffers "
                 "already grown) of level_search_kernel on the bench workload (bank 10k ops / 32 clients, tau_think 0, "
                 "Knossos-exact space)",
       "table_probe_algo_GBps": 586.0,
       "table_probe_source": "profiles/r2_probe_ceiling.md: 36.6 G random 16 B requests/s to an HBM-resident table"}
json.dump(out, open(os.path.join(ROOT, "profiles", "bench_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
