"""A/B of kernel experiment switches (env vars are read per call by the library)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jepsen_tigerbeetle_b200 import native, synth, history as H
m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
cases = [("think20 eager", 20, True), ("think5 eager", 5, True), ("think5 exact", 5, False)]
hs = {t: synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=t * 1e6, stale_read=True)) for t in (20, 5)}
for chain in (0, 1, 4, 16):
    for narrow in (0, 1):
        os.environ["JTB_CHAIN"] = str(chain); os.environ["JTB_NARROW_CAS"] = str(narrow)
        row = []
        for name, t, eager in cases:
            with native.Context(eager_reads=eager) as ctx:
                best = min(ctx.check_linearizable(hs[t], m)["seconds_kernel"] for _ in range(3))
            row.append(f"{name} {best*1e3:.1f} ms")
        print(f"chain={chain} narrow_cas={narrow}: " + " | ".join(row), flush=True)
