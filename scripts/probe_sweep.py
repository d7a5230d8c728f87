"""Judge item: prove or retract the random-probe ceiling.  Random 16 B gathers (the search kernel's probe instruction)
over footprints 64 MiB .. 16 GiB x loads in flight per thread {4, 8, 16} x {16 B slot, 32 B sector}; plus the keyed
table bench (insert + probe with real keys) at 3 table sizes.  Writes gpurun_out/probe_sweep.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jepsen_tigerbeetle_b200 import native

out = {"gather": [], "table_bench": []}
sizes_mib = [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
if len(sys.argv) > 1:
    sizes_mib = [int(x) for x in sys.argv[1].split(",")]
with native.Context(device=0) as ctx:
    for mib in sizes_mib:
        for wide in (1, 2):
            for u in (4, 8, 16):
                for ctas in (4, 8):
                    r = native.gather_bench(ctx, mib << 20, in_flight=u, wide=wide, iters=256, ctas_per_sm=ctas, rounds=3)
                    r["table_MiB"] = mib
                    out["gather"].append(r)
                    print(json.dumps(r), flush=True)
for mib in (256, 2048, 8192):
    with native.Context(device=0, table_bytes=mib << 20) as ctx:
        n = (mib << 20) // 16 // 4      # load 0.25
        for variant in (0, 2):
            r = ctx.table_bench(n, variant, rounds=3)
            r["table_MiB"] = mib
            r["probe_Gps"] = n * r["rounds"] / r["probe_seconds"] / 1e9
            r["insert_Gps"] = n / r["insert_seconds"] / 1e9
            out["table_bench"].append(r)
            print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_sweep.json"), "w"), indent=1)
