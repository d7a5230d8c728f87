"""One level-engine search (bank 10k ops / 32 clients) for profiling: python scripts/prof_level.py THINK_MS exact|eager [reps] [engine]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jepsen_tigerbeetle_b200 import native, synth, history as H
think = float(sys.argv[1]) if len(sys.argv) > 1 else 5
exact = len(sys.argv) > 2 and sys.argv[2] == "exact"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
engine = sys.argv[4] if len(sys.argv) > 4 else "level"
h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=think * 1e6))
m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
with native.Context(device=0, eager_reads=not exact, engine=engine) as ctx:
    for _ in range(reps):
        r = ctx.check_linearizable(h, m)
    r.pop("shards")
    r["stats"] = ctx.stats()
print(json.dumps(r))
