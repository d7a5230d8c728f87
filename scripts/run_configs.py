"""Run the five BASELINE.json configs on the GPU and on the CPU oracle; write gpurun_out/configs_r1.json.
CPU-oracle runs that would take minutes are bounded by a config budget and reported as rates."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jepsen_tigerbeetle_b200 import native, synth, history as H
import oracle

oracle.build()
SECTIONS = set(sys.argv[1:]) or {"c1", "c2", "c3", "c4", "c5"}
TAG = "_".join(sorted(SECTIONS))
MAX_THREADS = 16   # bound host memory: every oracle thread owns a visited set
ctx = native.Context(device=0, table_bytes=0)
out = {"host_cores": os.cpu_count(), "sections": sorted(SECTIONS)}
V = {0: "valid", 1: "unknown", 2: "invalid"}


def lin(name, h, m, cpu_budget=0, threads=1, reps=2):
    g = None
    for _ in range(reps):
        g = ctx.check_linearizable(h, m)
    t = time.perf_counter()
    o = oracle.check_linearizable(h, m, oracle.ALGO_WGL_COMPACT, max_configs=cpu_budget, n_threads=threads)
    cpu_s = time.perf_counter() - t
    rec = {"gpu_verdict": V[g["valid"]], "cpu_verdict": V[o["valid"]],
           "gpu_witness": [s["witness_index"] for s in g["shards"]][:4], "cpu_witness": [s["witness_index"] for s in o["shards"]][:4],
           "gpu_failures": g["n_failures"], "cpu_failures": o["n_failures"],
           "gpu_configs": g["configs"], "cpu_configs": o["configs"], "gpu_probes": g["probes"],
           "gpu_kernel_s": g["seconds_kernel"], "gpu_total_s": g["seconds_total"], "cpu_s": cpu_s,
           "cpu_threads": threads, "cpu_budget": cpu_budget, "key_bytes": g["key_bytes"],
           "gpu_configs_per_s": g["configs"] / max(g["seconds_kernel"], 1e-9), "cpu_configs_per_s": o["configs"] / max(cpu_s, 1e-9),
           "algo_GBps": g["hbm_bytes_algorithmic"] / max(g["seconds_kernel"], 1e-9) / 1e9}
    if cpu_budget == 0 or o["valid"] != 1:
        rec["speedup_time_to_verdict"] = cpu_s / g["seconds_total"]
        same = (g["valid"] == o["valid"] and [s["valid"] for s in g["shards"]] == [s["valid"] for s in o["shards"]]
                and [s["witness_index"] for s in g["shards"]] == [s["witness_index"] for s in o["shards"]])
        rec["parity"] = bool(same)
    out[name] = rec
    print(name, json.dumps(rec), flush=True)


# C1: set-full, 100 ops, 4 clients
for seed in ((1, 2, 3) if "c1" in SECTIONS else ()):
    h = synth.config_c1(seed=seed)
    g, o = ctx.check_set_full(h, True), oracle.check_set_full(h, True)
    out[f"c1_seed{seed}"] = {"gpu_verdict": V[g["valid"]], "cpu_verdict": V[o["valid"]], "parity": g["shards"] == o["shards"],
                            "gpu_total_s": g["seconds"], "cpu_s": o["seconds"], "shard": g["shards"][0]}
    print("c1", seed, out[f"c1_seed{seed}"], flush=True)
# C2: 1k-op cas-register, 16 clients
mc = H.make_model(H.MODEL_CAS_REGISTER)
for seed in ((1, 2, 3) if "c2" in SECTIONS else ()):
    for p_info in (0.0, 0.05):
        for stale in (False, True):
            lin(f"c2_seed{seed}_pinfo{p_info}_{'stale' if stale else 'valid'}", synth.config_c2(seed=seed, p_info=p_info, stale_read=stale), mc)
# C3: 10k-op bank, 32 clients (tau_think 5 ms); CPU bounded to 20 M configs except seed 1 (measured in full separately)
mb = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
for seed in ((1, 2, 3) if "c3" in SECTIONS else ()):
    for stale in (False, True):
        h = synth.generate(synth.SynthSpec("bank", 10000, 32, seed, tau_think_ns=5e6, stale_read=stale))
        lin(f"c3_seed{seed}_{'stale' if stale else 'valid'}", h, mb, cpu_budget=20_000_000)
if "c3" in SECTIONS:
    h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=5e6, p_info=0.02))
    lin("c3_seed1_pinfo0.02_valid", h, mb, cpu_budget=20_000_000)
# C4: set-full, 100k ops, 64 clients, K = 64 and K = 8 ledgers: exact set-full scan and WGL set model
ms = H.make_model(H.MODEL_SET)
for K in ([64] if "c4" in SECTIONS else []) + ([8] if "c4k8" in SECTIONS else []):
    h = synth.config_c4(seed=1, n_keys=K)
    for _ in range(2):
        g = ctx.check_set_full(h, True)
    o = oracle.check_set_full(h, True)
    out[f"c4_K{K}_setfull"] = {"gpu_verdict": V[g["valid"]], "cpu_verdict": V[o["valid"]], "parity": g["shards"] == o["shards"],
                               "gpu_kernel_s": g["seconds_kernel"], "gpu_total_s": g["seconds"], "cpu_s": o["seconds"],
                               "payload_ints": int(h.payload.shape[0]), "events": h.n_events}
    print(f"c4_K{K}_setfull", out[f"c4_K{K}_setfull"], flush=True)
    lin(f"c4_K{K}_wgl_set", h, ms, threads=min(K, MAX_THREADS), cpu_budget=20_000_000)
# C5: 50k-op adversarial cas-register, 30% :info, K = 256 keys, 8 clients per key
if "c5" in SECTIONS:
    h = synth.config_c5(seed=1)
    lin("c5_K256", h, mc, threads=MAX_THREADS, cpu_budget=5_000_000)
    h = synth.config_c5(seed=1, stale_read=True, n_values=30)
    lin("c5_K256_stale", h, mc, threads=MAX_THREADS, cpu_budget=5_000_000)
if "monster" in SECTIONS:
    ctx.close()
    ctx = native.Context(device=0, max_configs=400_000_000, time_budget_ms=8000, table_bytes=16 << 30)
    h = synth.config_c5(seed=1, n_keys=8, n_ops=50000)
    lin("c5_K8_monster_budget8s", h, mc, threads=8, cpu_budget=20_000_000, reps=1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/configs_r1_{TAG}.json", "w"), indent=1)
print("done")
