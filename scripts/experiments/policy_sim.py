"""Driver for policy_sim.cpp (design experiment, see its header): dumps synthetic register histories and prints the
number of configs each queue policy inserts before it finds the linearization.  CPU only.
usage: python scripts/experiments/policy_sim.py [W] [MAX_CONFIGS]"""
import json, os, struct, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from jepsen_tigerbeetle_b200 import synth, history as H

HERE = os.path.dirname(os.path.abspath(__file__))
exe = os.path.join(tempfile.gettempdir(), "jtb_policy_sim")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"),
                       os.path.join(HERE, "policy_sim.cpp"), "-o", exe])
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
MAXC = int(sys.argv[2]) if len(sys.argv) > 2 else 3_000_000
specs = [synth.SynthSpec('cas-register', 2500, 24, 809007372, p_info=0.3, tau_think_ns=20e6, n_values=30),
         synth.SynthSpec('register', 1000, 24, 902980068, p_info=0.3, tau_think_ns=5e6, n_values=30, stale_read=True),
         synth.SynthSpec('register', 2500, 40, 321354213, p_info=0.1, tau_think_ns=20e6, n_values=30, stale_by=3),
         synth.SynthSpec('cas-register', 1000, 16, 1, p_info=0.05),
         synth.SynthSpec('cas-register', 1000, 16, 1, p_info=0.0)]
KIND = {"register": H.MODEL_REGISTER, "cas-register": H.MODEL_CAS_REGISTER}
for sp in specs:
    h = synth.generate(sp)
    m = H.make_model(KIND[sp.model])
    o = oracle.check_linearizable(h, m, 3, eager_reads=True, max_configs=5_000_000)
    path = os.path.join(tempfile.gettempdir(), "jtb_policy_hist.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("<qii", len(h.type), KIND[sp.model], int(m.init_value)))
        for arr, dt in ((h.type, np.uint8), (h.f, np.uint8), (h.process, np.int32), (h.index, np.int32), (h.a, np.int32), (h.b, np.int32)):
            f.write(np.ascontiguousarray(arr, dtype=dt).tobytes())
    print(f"{sp.model} {sp.n_ops} ops / {sp.n_clients} clients, p_info {sp.p_info}: CPU depth-first (knossos order) = {o['configs']} configs, verdict {o['valid']}", flush=True)
    for policy in ("fifo", "lifo", "rank", "rank-crash"):
        for w in sorted({1, 64, W}):
            r = subprocess.run([exe, path, policy, str(w), str(MAXC)], capture_output=True, text=True)
            print("   ", r.stdout.strip() or r.stderr.strip()[-200:], flush=True)
