"""One search (bank, 32 clients) for profiling under ncu: python scripts/prof_case.py [think_ms] [stale] [reps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jepsen_tigerbeetle_b200 import native, synth, history as H
think = float(sys.argv[1]) if len(sys.argv) > 1 else 10
stale = len(sys.argv) > 2 and sys.argv[2] == "stale"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
exact = len(sys.argv) > 4 and sys.argv[4] == "exact"   # Knossos-exact space (no eager reads)
h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=think * 1e6, stale_read=stale))
m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
ctx = native.Context(device=0, eager_reads=not exact)
for _ in range(reps):
    r = ctx.check_linearizable(h, m)
r.pop("shards")
r["stats"] = ctx.stats()
print(json.dumps(r))
