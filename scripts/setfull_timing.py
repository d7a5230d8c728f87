"""set-full e2e vs kernel time on BASELINE config #4 (100k ops): pageable vs page-locked id lists, warm context."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jepsen_tigerbeetle_b200 import native, synth
out = {}
for K in (64, 8):
    h = synth.config_c4(seed=1, n_keys=K)
    hp = native.pin_history(h)
    with native.Context() as ctx:
        for name, hh in (("pageable", h), ("pinned", hp)):
            best = None
            for _ in range(4):
                t = time.perf_counter(); r = ctx.check_set_full(hh, True); dt = time.perf_counter() - t
                if best is None or dt < best[0]:
                    best = (dt, r["seconds_kernel"], r["seconds"])
            out[f"K{K}_{name}"] = {"call_ms": 1e3 * best[0], "native_total_ms": 1e3 * best[2], "kernel_ms": 1e3 * best[1],
                                  "payload_MB": h.payload.nbytes / 1e6, "valid": r["valid"]}
            print(f"K{K}_{name}", json.dumps(out[f"K{K}_{name}"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "setfull_timing.json"), "w"), indent=1)
