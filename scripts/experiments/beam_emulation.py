"""Runner of beam_emulation.cpp (see its header): python scripts/experiments/beam_emulation.py"""
import ctypes as C, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from jepsen_tigerbeetle_b200 import synth, history as H
from jepsen_tigerbeetle_b200.history import as_c_history
so = "/tmp/libbeam.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + ROOT, "-DCMAX=15", "-o", so,
                       os.path.join(ROOT, "scripts/experiments/beam_emulation.cpp"), os.path.join(ROOT, "jepsen_tigerbeetle_b200/csrc/jtb_prep.cpp")])
L = C.CDLL(so)
M = {"register": H.MODEL_REGISTER, "cas-register": H.MODEL_CAS_REGISTER, "bank": H.MODEL_BANK}
specs = [synth.SynthSpec('cas-register', 2500, 24, 809007372, p_info=0.3, tau_think_ns=20e6, n_values=30),
         synth.SynthSpec('register', 1000, 24, 902980068, p_info=0.3, tau_think_ns=5e6, n_values=30, stale_read=True),
         synth.SynthSpec('register', 2500, 40, 321354213, p_info=0.1, tau_think_ns=20e6, n_values=30, stale_by=3),
         synth.SynthSpec('cas-register', 1000, 16, 1, p_info=0.05), synth.SynthSpec('cas-register', 6250, 8, 1, p_info=0.3),
         synth.SynthSpec('bank', 2000, 32, 1, tau_think_ns=5e6, p_info=0.02), synth.SynthSpec('bank', 10000, 32, 1, tau_think_ns=20e6, p_info=0.02)]
for sp in specs:
    h = synth.generate(sp)
    m = H.make_model(M[sp.model], accounts=range(1, 9)) if sp.model == "bank" else H.make_model(M[sp.model])
    ch = as_c_history(h)
    for pol in (1, 2):
        for W in (256, 1024, 4096, 16384):
            out = (C.c_ulonglong * 5)()
            t = time.time()
            L.beam_run(C.byref(ch), C.byref(m), W, pol, 1, out)
            print(sp.model, sp.n_ops, sp.n_clients, sp.p_info, "policy", pol, "W", W, "found" if out[0] else "DIED/BUDGET(40M)", "configs", out[1],
                  "levels", out[2], "max width", out[3], f"{time.time() - t:.1f}s", flush=True)
            if out[0]:
                break
