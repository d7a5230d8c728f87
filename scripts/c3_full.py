"""SURVEY 8(d) C3 exactly as written: bank 10k ops / 32 clients, tau_think 0, seeds 1-3, valid + one stale read
(Knossos-exact space), and the p_info 0.02 variant (tau_think 5 ms and 0).  Prints verdict, witness, time."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jepsen_tigerbeetle_b200 import native, synth, history as H
m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
out = []
V = {0: "valid", 1: "unknown", 2: "invalid"}
with native.Context(eager_reads=False, time_budget_ms=60_000) as ctx:
    for seed in (1, 2, 3):
        for stale in (False, True):
            h = synth.generate(synth.SynthSpec("bank", 10000, 32, seed, tau_think_ns=0, stale_read=stale))
            r = ctx.check_linearizable(h, m)
            rec = {"case": f"think0 seed{seed} {'stale' if stale else 'valid'} exact", "verdict": V[r["valid"]], "configs": r["configs"],
                   "witness_index": r["shards"][0]["witness_index"], "mutated_op_index": h.meta.get("mutated_op_index"),
                   "seconds": r["seconds_total"], "kernel_s": r["seconds_kernel"]}
            out.append(rec); print(json.dumps(rec), flush=True)
with native.Context(time_budget_ms=60_000) as ctx:
    for think in (5, 0):
        for seed in (1, 2):
            h = synth.generate(synth.SynthSpec("bank", 10000, 32, seed, tau_think_ns=think * 1e6, p_info=0.02))
            r = ctx.check_linearizable(h, m)
            st = ctx.stats()
            rec = {"case": f"p_info 0.02 think{think} seed{seed} (eager default)", "verdict": V[r["valid"]], "cause": r["shards"][0]["cause"],
                   "configs": r["configs"], "seconds": r["seconds_total"], "beam_attempts": st["beam_attempts"], "beam_decided": st["beam_decided"],
                   "beam_levels": st["beam_levels"], "key_bytes": r["key_bytes"]}
            out.append(rec); print(json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "c3_full.json"), "w"), indent=1)
