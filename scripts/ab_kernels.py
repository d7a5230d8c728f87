"""A/B on one box: warp-per-config kernel (jtb_wgl.cuh) vs thread-per-config kernel (jtb_search.cuh), each in a fresh
process (the kernel is chosen by env JTB_KERNEL), on the bench workload (bank 10k ops / 32 clients, tau_think 5 ms)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import json, os, sys
sys.path.insert(0, %r)
from jepsen_tigerbeetle_b200 import native, synth, history as H
m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
mc = H.make_model(H.MODEL_CAS_REGISTER)
res = {}
for eager in (False, True):
    with native.Context(eager_reads=eager) as ctx:
        for stale in (False, True):
            h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=5e6, stale_read=stale))
            best = None
            for rep in range(3):
                r = ctx.check_linearizable(h, m)
                if best is None or r["seconds_kernel"] < best["seconds_kernel"]:
                    best = r
            st = ctx.stats()
            res["bank_%%s_%%s" %% ("eager" if eager else "exact", "stale" if stale else "valid")] = {
                "valid": best["valid"], "configs": best["configs"], "probes": best["probes"],
                "kernel_ms": 1e3 * best["seconds_kernel"], "total_ms": 1e3 * best["seconds_total"],
                "Mconfigs_s": best["configs"] / best["seconds_kernel"] / 1e6, "grid": st["grid"],
                "attempts": st["attempts"], "polls": st["idle_polls"], "max_probe": st["max_probe_len"]}
with native.Context() as ctx:
    for p_info in (0.0, 0.05):
        h = synth.config_c2(seed=1, p_info=p_info)
        best = min((ctx.check_linearizable(h, mc) for _ in range(3)), key=lambda r: r["seconds_total"])
        res["c2_pinfo%%g" %% p_info] = {"valid": best["valid"], "configs": best["configs"],
                                       "kernel_ms": 1e3 * best["seconds_kernel"], "total_ms": 1e3 * best["seconds_total"]}
print(json.dumps(res))
''' % ROOT
out = {}
variants = [("warp", {"JTB_KERNEL": "warp"}), ("tpc", {"JTB_KERNEL": "tpc"})]
for extra in sys.argv[1:]:          # e.g. tpc:JTB_CTAS_PER_SM=2
    name, kv = extra.split(":", 1)
    env = {"JTB_KERNEL": "tpc"}
    for item in kv.split(","):
        k, v = item.split("=")
        env[k] = v
    variants.append((name, env))
for name, envx in variants:
    env = dict(os.environ); env.update(envx)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    try:
        out[name] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:  # noqa: BLE001
        out[name] = {"error": r.stderr[-2000:], "stdout": r.stdout[-500:]}
    print(name, json.dumps(out[name]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_kernels.json"), "w"), indent=1)
