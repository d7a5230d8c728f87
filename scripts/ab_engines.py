"""A/B on one box: level engine (jtb_level.cuh) vs work-list engine (warp kernel, jtb_wgl.cuh) on the bench workload
(bank 10k ops / 32 clients, tau_think 5 ms) in both search spaces, C2, and harder bank instances (smaller tau_think)
that only the level engine is expected to finish.  Writes gpurun_out/ab_engines.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jepsen_tigerbeetle_b200 import native, synth, history as H

m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
mc = H.make_model(H.MODEL_CAS_REGISTER)
out = {}
names = ["configs", "probes", "levels", "max_width", "narrow_levels", "x5", "max_probe", "max_window", "grid", "buf_cap", "attempts"]


def run(tag, ctx, h, model, reps=3):
    best = None
    for _ in range(reps):
        r = ctx.check_linearizable(h, model)
        if best is None or r["seconds_kernel"] < best["seconds_kernel"]:
            best = r
    st = ctx.stats()
    rec = {"valid": best["valid"], "configs": best["configs"], "probes": best["probes"],
           "kernel_ms": 1e3 * best["seconds_kernel"], "total_ms": 1e3 * best["seconds_total"],
           "Mconfigs_s": best["configs"] / max(best["seconds_kernel"], 1e-9) / 1e6, "cause": best["shards"][0]["cause"],
           "stats": {k: st[k] for k in ("expansions", "ring_tail", "ring_head", "max_probe_len", "table_slots", "grid", "attempts")}}
    out[tag] = rec
    print(tag, json.dumps(rec), flush=True)


which = sys.argv[1:] or ["base", "hard"]
if "base" in which:
    for engine in ("worklist", "level"):
        for eager in (False, True):
            with native.Context(eager_reads=eager, engine=engine) as ctx:
                for stale in (False, True):
                    h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=5e6, stale_read=stale))
                    run(f"{engine}_bank_think5_{'eager' if eager else 'exact'}_{'stale' if stale else 'valid'}", ctx, h, m)
        with native.Context(engine=engine) as ctx:
            for p_info in (0.0, 0.05):
                run(f"{engine}_c2_pinfo{p_info}", ctx, synth.config_c2(seed=1, p_info=p_info), mc)
if "hard" in which:
    for think_ms, eager in ((3, False), (2, False), (0, True), (1, False), (0, False)):
        with native.Context(eager_reads=eager, engine="level", time_budget_ms=60_000) as ctx:
            h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=think_ms * 1e6))
            run(f"level_bank_think{think_ms}_{'eager' if eager else 'exact'}_valid", ctx, h, m, reps=1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_engines.json"), "w"), indent=1)
