"""First-contact GPU exploration: search throughput on C3-shaped histories + the K2 table bench."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jepsen_tigerbeetle_b200 import native, synth, history as H

out = {}
ctx = native.Context(device=0)
m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
for think_ms, stale in ((20, False), (20, True), (10, False), (10, True), (5, False), (5, True)):
    h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=think_ms * 1e6, stale_read=stale))
    for rep in range(2):
        r = ctx.check_linearizable(h, m)
    key = f"bank_think{think_ms}_{'stale' if stale else 'valid'}"
    r.pop("shards")
    r["stats"] = ctx.stats()
    r["configs_per_s"] = r["configs"] / r["seconds_kernel"]
    r["probes_per_s"] = r["probes"] / r["seconds_kernel"]
    r["algo_GBps"] = r["hbm_bytes_algorithmic"] / r["seconds_kernel"] / 1e9
    out[key] = r
    print(key, json.dumps(r), flush=True)
ctx.close()
ctx = native.Context(device=0, table_bytes=8 << 30)
for variant in (0, 1, 2):
    n = 1 << 27  # 134M keys in 512M slots (load 0.25)
    r = ctx.table_bench(n, variant, rounds=3)
    r["probe_Gps"] = n * r["rounds"] / r["probe_seconds"] / 1e9
    r["insert_Gps"] = n / r["insert_seconds"] / 1e9
    r["probe_algo_GBps"] = 16 * r["probe_Gps"]
    out[f"table_v{variant}"] = r
    print("table", variant, json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/explore1.json", "w"), indent=1)
