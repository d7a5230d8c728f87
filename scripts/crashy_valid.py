"""The three crash-heavy VALID histories the soak test found (GPU :unknown vs CPU valid) + a few more."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jepsen_tigerbeetle_b200 import native, synth, history as H
import oracle
specs = [synth.SynthSpec('cas-register', 2500, 24, 809007372, p_info=0.3, tau_think_ns=20e6, n_values=30),
         synth.SynthSpec('register', 1000, 24, 902980068, p_info=0.3, tau_think_ns=5e6, n_values=30, stale_read=True),
         synth.SynthSpec('register', 2500, 40, 321354213, p_info=0.1, tau_think_ns=20e6, n_values=30, stale_by=3),
         synth.SynthSpec('cas-register', 1000, 16, 1, p_info=0.05), synth.SynthSpec('cas-register', 50000, 2048, 1, p_info=0.3, n_keys=256, grouped_keys=True),
         synth.SynthSpec('cas-register', 50000, 64, 1, p_info=0.3, n_keys=8, grouped_keys=True)]
rows = []
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M = {"register": H.MODEL_REGISTER, "cas-register": H.MODEL_CAS_REGISTER}
with native.Context(max_configs=400_000_000) as ctx:
    for sp in specs:
        h = synth.generate(sp)
        m = H.make_model(M[sp.model])
        t = time.perf_counter(); o = oracle.check_linearizable(h, m, 3, eager_reads=True, n_threads=8, max_configs=5_000_000); tc = time.perf_counter() - t
        for _ in range(REPS):      # the last (warm) call is reported
            g = ctx.check_linearizable(h, m)
        st = ctx.stats()
        print("   beam levels", st["beam_levels"], "configs", st["beam_configs"], "decided", st["beam_decided"], "attempts", st["beam_attempts"], flush=True)
        print("   scouts", st["scouts"], "steps", st["scout_steps"], "configs", st["scout_configs"], "decided", st["scout_decided"], flush=True)
        rows.append({"model": sp.model, "n_ops": sp.n_ops, "n_clients": sp.n_clients, "p_info": sp.p_info, "n_keys": sp.n_keys,
                     "gpu_valid": g["valid"], "gpu_configs": g["configs"], "gpu_seconds": g["seconds_total"],
                     "beam_levels": st["beam_levels"], "beam_configs": st["beam_configs"], "beam_decided": st["beam_decided"], "beam_attempts": st["beam_attempts"],
                     "scout_steps": st["scout_steps"], "scout_configs": st["scout_configs"], "scout_decided": st["scout_decided"],
                     "cpu_valid": o["valid"], "cpu_configs": o["configs"], "cpu_seconds": tc})
        print(sp.model, sp.n_ops, sp.n_clients, sp.p_info, "keys", sp.n_keys, "| gpu", g["valid"], g["configs"], round(g["seconds_total"] * 1e3, 1), "ms | cpu", o["valid"], o["configs"], round(tc * 1e3, 1), "ms", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/crashy_valid.json", "w"), indent=1)
