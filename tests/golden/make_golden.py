"""Generates tests/golden/*.json: known-answer vectors for seeded synthetic histories.

The reference holds NO golden vectors for this path and cannot run here (JVM), so these are produced
by the CPU oracle (4 cross-checked deciders, see oracle/oracle_common.h) — they pin the GPU path and
guard the oracle against regressions; they do not pin the oracle to the reference.
Run from the repo root:  python tests/golden/make_golden.py"""
import dataclasses
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from jepsen_tigerbeetle_b200 import history as H, synth  # noqa: E402
import oracle  # noqa: E402

MODEL = {"register": H.MODEL_REGISTER, "cas-register": H.MODEL_CAS_REGISTER, "set": H.MODEL_SET, "bank": H.MODEL_BANK}


def model_of(name):
    return H.make_model(MODEL[name], accounts=range(1, 9)) if name == "bank" else H.make_model(MODEL[name])


LIN_SPECS = []
for model in ("register", "cas-register", "bank", "set"):
    for seed in range(1, 7):
        LIN_SPECS.append(dict(model=model, n_ops=400, n_clients=6, seed=seed, p_info=0.05 if seed % 2 else 0.0,
                              stale_read=seed % 3 != 0, tau_think_ns=5e6, n_values=4, n_keys=1 if seed < 5 else 3))
LIN_SPECS.append(dict(model="cas-register", n_ops=1000, n_clients=16, seed=1, n_values=30, stale_read=True))
LIN_SPECS.append(dict(model="bank", n_ops=3000, n_clients=32, seed=2, tau_think_ns=20e6, stale_read=True))
LIN_SPECS.append(dict(model="bank", n_ops=3000, n_clients=32, seed=3, tau_think_ns=20e6))

SF_SPECS = [dict(model="set", n_ops=100, n_clients=4, seed=s, final_reads=True) for s in (1, 2, 3)]
SF_SPECS += [dict(model="set", n_ops=3000, n_clients=16, seed=s, n_keys=4, p_info=0.02, final_reads=True) for s in (4, 5)]


def main():
    lin = []
    for spec in LIN_SPECS:
        h = synth.generate(synth.SynthSpec(**spec))
        m = model_of(spec["model"])
        r = oracle.check_linearizable(h, m, oracle.ALGO_WGL_COMPACT)
        r2 = oracle.check_linearizable(h, m, oracle.ALGO_WGL)
        assert [(s["valid"], s["witness_index"]) for s in r["shards"]] == [(s["valid"], s["witness_index"]) for s in r2["shards"]]
        re_ = oracle.check_linearizable(h, m, oracle.ALGO_WGL_COMPACT, eager_reads=True)
        assert [(s["valid"], s["witness_index"]) for s in r["shards"]] == [(s["valid"], s["witness_index"]) for s in re_["shards"]]
        exhaustive = r["valid"] == H.INVALID and h.n_shards == 1
        lin.append({"spec": spec, "valid": r["valid"],
                    "shards": [{k: s[k] for k in ("valid", "witness_index", "previous_ok_index")} for s in r["shards"]],
                    "configs_if_exhaustive": r["configs"] if exhaustive else None,          # knossos-exact space
                    "configs_if_exhaustive_eager": re_["configs"] if exhaustive else None})  # with eager reads
    sf = []
    for spec in SF_SPECS:
        h = synth.generate(synth.SynthSpec(**spec))
        for linz in (True, False):
            r = oracle.check_set_full(h, linz)
            sf.append({"spec": spec, "linearizable": linz, "valid": r["valid"], "shards": r["shards"],
                       "elem_outcome": [int(x) for x in r["elem_outcome"]], "elem_latency_ms": [int(x) for x in r["elem_latency_ms"]]})
    here = os.path.dirname(os.path.abspath(__file__))
    json.dump(lin, open(os.path.join(here, "lin_golden.json"), "w"), indent=0)
    json.dump(sf, open(os.path.join(here, "setfull_golden.json"), "w"), indent=0)
    print(len(lin), "linearizability vectors,", len(sf), "set-full vectors")


if __name__ == "__main__":
    main()
