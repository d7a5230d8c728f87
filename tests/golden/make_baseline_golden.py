"""Generates tests/golden/baseline_golden.json: the CPU oracle run TO COMPLETION on BASELINE.json's own configs
(SURVEY.md §8(d)): the exact bench instance (C3 bank 10k ops / 32 clients, tau_think 5 ms, seeds 1-3, valid and
one-stale-read, Knossos-exact and eager-read spaces), full-size C4 (100k-op set-full, K = 64 and K = 8: set-full scan
and the WGL search with the grow-only-set model) and full-size C5 (50k ops, 30 % :info, K = 256; K = 8 "monster" under
a configuration budget), plus the poisoned C5 / C4 histories that bench.py shards over the GPUs.

The reference holds no golden vectors for this path and cannot run here (JVM, un-vendored Knossos): these records are
produced by oracle/ (DESIGN.md section 2 - "parity unpinned") and pin the GPU path at BASELINE sizes, where running
the oracle inside the test would take minutes per case.  `python tests/golden/make_baseline_golden.py [case-prefix]`
from the repo root; takes ~10 minutes on 8 cores (the C3 Knossos-exact searches are 30-90 s each, single-threaded like
knossos.wgl).  Config counts of VALID histories are the depth-first count of the CPU port (informational: the GPU
visits a different part of the space); for INVALID histories the search is exhaustive and the count is a parity key."""
import hashlib
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402

from jepsen_tigerbeetle_b200 import history as H, synth  # noqa: E402
import baseline_cases as BC  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "baseline_golden.json")


def sha(*arrays):
    d = hashlib.sha256()
    for a in arrays:
        d.update(np.ascontiguousarray(a).tobytes())
    return d.hexdigest()[:32]


def run_case(name):
    import oracle
    case = BC.CASES[name]
    h = BC.build_history(case)
    t0 = time.perf_counter()
    rec = {"case": name, "kind": case["kind"], "events": int(h.n_events), "keys": int(h.n_shards)}
    if case["kind"] == "lin":
        m = BC.model_of(case["model"])
        r = oracle.check_linearizable(h, m, case.get("oracle_algo", oracle.ALGO_WGL_COMPACT), max_configs=case.get("max_configs", 0),
                                      n_threads=case.get("threads", 1), eager_reads=case.get("eager", False))
        rec.update({"valid": r["valid"], "n_failures": r["n_failures"], "configs": r["configs"],
                    "shard_valid": [s["valid"] for s in r["shards"]],
                    "shard_witness": [s["witness_index"] for s in r["shards"]],
                    "shard_previous_ok": [s["previous_ok_index"] for s in r["shards"]],
                    "shard_configs": [s["configs"] for s in r["shards"]],
                    "shard_cause": [s["cause"] for s in r["shards"]],
                    "max_configs": case.get("max_configs", 0), "eager": bool(case.get("eager", False)),
                    "oracle_algo": case.get("oracle_algo", 3), "compare_counts": bool(case.get("compare_counts", True))})
    else:
        r = oracle.check_set_full(h, True)
        rec.update({"valid": r["valid"], "n_failures": r["n_failures"], "shards": r["shards"],
                    "elem_outcome_sha": sha(np.asarray(r["elem_outcome"], dtype=np.int32)),
                    "elem_latency_sha": sha(np.asarray(r["elem_latency_ms"], dtype=np.int64)),
                    "n_elems": len(r["elem_outcome"])})
    rec["oracle_seconds"] = time.perf_counter() - t0
    return rec


def main():
    import oracle
    oracle.build()
    prefix = sys.argv[1] if len(sys.argv) > 1 else ""
    names = [n for n in BC.CASES if n.startswith(prefix)]
    old = {}
    if os.path.exists(OUT):
        old = {r["case"]: r for r in json.load(open(OUT))}
    # multi-threaded cases one at a time, single-threaded ones in a pool
    multi = [n for n in names if BC.CASES[n].get("threads", 1) > 1]
    single = [n for n in names if n not in multi]
    for n in multi:
        old[n] = run_case(n)
        print(n, old[n]["valid"], f'{old[n]["oracle_seconds"]:.1f}s', flush=True)
    with ProcessPoolExecutor(max_workers=int(os.environ.get("JTB_GOLDEN_WORKERS", "4"))) as pool:
        for rec in pool.map(run_case, single):
            old[rec["case"]] = rec
            print(rec["case"], rec["valid"], rec.get("configs"), f'{rec["oracle_seconds"]:.1f}s', flush=True)
    json.dump([old[k] for k in sorted(old)], open(OUT, "w"), indent=0)


if __name__ == "__main__":
    main()
