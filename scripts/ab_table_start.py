"""A/B: initial visited-table size (L2-resident 64 MiB vs 1 GiB).  Fresh context per measurement = cold start."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, time
sys.path.insert(0, %r)
from jepsen_tigerbeetle_b200 import native, synth, history as H
mb = H.make_model(H.MODEL_BANK, accounts=range(1, 9)); mc = H.make_model(H.MODEL_CAS_REGISTER)
cases = [("c2", synth.config_c2(seed=1), mc), ("c2info", synth.config_c2(seed=1, p_info=0.05), mc),
         ("b20", synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=20e6, stale_read=True)), mb),
         ("b5", synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=5e6, stale_read=True)), mb)]
out = []
for name, h, m in cases:
    cold = []
    for _ in range(2):
        with native.Context() as ctx:          # cold context: first call decides the start size
            r = ctx.check_linearizable(h, m); cold.append(r["seconds_total"])
            warm = min(ctx.check_linearizable(h, m)["seconds_total"] for _ in range(3))
    out.append("%%s cold %%.1f warm %%.1f ms" %% (name, min(cold) * 1e3, warm * 1e3))
with native.Context(eager_reads=False) as ctx:
    h = cases[3][1]
    r = ctx.check_linearizable(h, mb); c = r["seconds_total"]
    w = min(ctx.check_linearizable(h, mb)["seconds_total"] for _ in range(2))
    out.append("b5exact cold %%.1f warm %%.1f ms" %% (c * 1e3, w * 1e3))
print(" | ".join(out))
''' % ROOT
for rep in range(2):
    for mbs in ("1024", "64", "256"):
        env = dict(os.environ); env["JTB_TABLE_START_MB"] = mbs
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=250)
        print(f"rep{rep} start={mbs:>4s} MiB: {r.stdout.strip()} {r.stderr.strip()[-200:]}", flush=True)
