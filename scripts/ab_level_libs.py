"""A/B of library builds (level-engine launch shapes: -DJTB_LV_WARPS / -DJTB_LV_CTAS) inside ONE gpurun call.
Each build runs in a fresh process (JTB_LIB_PATH); extra args NAME:ENV=V,... add env variants of the default build."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import json, os, sys
sys.path.insert(0, %r)
from jepsen_tigerbeetle_b200 import native, synth, history as H
m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
mc = H.make_model(H.MODEL_CAS_REGISTER)
out = {}
def best(ctx, h, model, reps=3):
    r = min((ctx.check_linearizable(h, model) for _ in range(reps)), key=lambda r: r["seconds_kernel"])
    return {"ms": round(1e3 * r["seconds_kernel"], 2), "Gcfg_s": round(r["configs"] / r["seconds_kernel"] / 1e9, 3), "v": r["valid"]}
cases = os.environ.get("AB_CASES", "t5x,t5e,c2,t2x,t0x").split(",")
with native.Context(eager_reads=False, engine="level") as ctx:
    if "t5x" in cases: out["think5_exact"] = best(ctx, synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=5e6)), m)
    if "t2x" in cases: out["think2_exact"] = best(ctx, synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=2e6)), m, 2)
    if "t0x" in cases: out["think0_exact"] = best(ctx, synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=0)), m, 1)
with native.Context(engine="level") as ctx:
    if "t5e" in cases: out["think5_eager"] = best(ctx, synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=5e6)), m)
    if "c2" in cases: out["c2"] = best(ctx, synth.config_c2(seed=1), mc)
print(json.dumps(out))
''' % ROOT
variants = [("default", {})]
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_ab_*.so"))):
    variants.append((os.path.basename(f)[10:-3], {"JTB_LIB_PATH": f}))
for extra in sys.argv[1:]:
    name, kv = extra.split(":", 1)
    variants.append((name, dict(item.split("=") for item in kv.split(","))))
res = {}
for name, envx in variants:
    env = dict(os.environ); env.update(envx)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    try:
        res[name] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:  # noqa: BLE001
        res[name] = {"error": r.stderr[-1500:]}
    print(name, json.dumps(res[name]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_level_libs.json"), "w"), indent=1)
