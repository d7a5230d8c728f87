"""A/B of library builds inside ONE gpurun call (box-to-box variance is larger than the effects)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import glob
libs = {"default": ""}
libs.update({os.path.basename(f)[10:-3]: f for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_ab_*.so")))})
code = r'''
import os, sys
sys.path.insert(0, %r)
from jepsen_tigerbeetle_b200 import native, synth, history as H
m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
out = []
for t, eager in ((20, True), (5, True), (5, False)):
    h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=t * 1e6, stale_read=True))
    with native.Context(eager_reads=eager) as ctx:
        best = min(ctx.check_linearizable(h, m)["seconds_kernel"] for _ in range(3))
    out.append("%%d/%%s %%.1f ms" %% (t, "eager" if eager else "exact", best * 1e3))
print(" | ".join(out))
''' % ROOT
for rep in range(2):
    for name, path in libs.items():
        env = dict(os.environ)
        if path:
            env["JTB_LIB_PATH"] = path
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(f"rep{rep} {name:10s}: {r.stdout.strip()} {r.stderr.strip()[-200:]}", flush=True)
