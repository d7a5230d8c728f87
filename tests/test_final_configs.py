"""knossos' :configs of an INVALID verdict: the configurations stuck at the witness.
CPU tier: the oracle's restatement on known answers + structural properties.  The GPU tier (test_gpu_parity.py)
compares `jtb_final_configs` with it config for config."""
import numpy as np
import pytest

from jepsen_tigerbeetle_b200 import history as H, synth

from arbitrary import arbitrary_history


def _op(p, t, f, v, i):
    return {"process": p, "type": t, "f": f, "value": v, "index": i, "time": i * 1000}


def test_known_answer_register(oracle_mod):
    """w1 ok; w3 in flight; read -> 2 can never be linearized.  Two configs are stuck at the read: register 1 with
    the write of 3 still pending, and register 3 with it linearized early (it returns after the read)."""
    ops = [_op(0, "invoke", "write", 1, 0), _op(0, "ok", "write", 1, 1), _op(1, "invoke", "write", 3, 2),
           _op(0, "invoke", "read", None, 3), _op(0, "ok", "read", 2, 4), _op(1, "ok", "write", 3, 5)]
    h = H.flatten_ops(ops, "register")
    m = H.make_model(H.MODEL_REGISTER)
    r = oracle_mod.check_linearizable(h, m, 3)
    assert r["valid"] == H.INVALID and r["shards"][0]["witness_index"] == 4
    f = oracle_mod.final_configs(h, m, 0, 10)
    zeros = [0] * 8
    assert f == {"total": 2, "configs": [
        {"state": 1, "balances": zeros, "pending": [2, 3], "linearized_open": [], "crashed_linearized": 0},
        {"state": 3, "balances": zeros, "pending": [3], "linearized_open": [2], "crashed_linearized": 0}]}


def test_first_op_stuck_reports_the_initial_configuration(oracle_mod):
    ops = [_op(0, "invoke", "read", None, 0), _op(0, "ok", "read", 7, 1)]
    h = H.flatten_ops(ops, "register")
    m = H.make_model(H.MODEL_REGISTER, init_value=0)
    f = oracle_mod.final_configs(h, m, 0, 10)
    assert f["total"] == 1 and f["configs"][0]["state"] == 0 and f["configs"][0]["pending"] == [0]


def test_valid_history_has_no_final_configs(oracle_mod):
    h = synth.generate(synth.SynthSpec("cas-register", 100, 4, 1))
    assert oracle_mod.final_configs(h, H.make_model(H.MODEL_CAS_REGISTER), 0, 10)["total"] == -1


@pytest.mark.parametrize("model", ["register", "cas-register", "bank"])
def test_structure_on_arbitrary_invalid_histories(oracle_mod, model):
    """Every final config has the witness' invocation pending, pending and linearized-open are disjoint, the list is
    in the canonical order, and a bank config's balances sum to the initial total (zero)."""
    m = H.make_model(H.MODEL_BANK, accounts=range(1, 4)) if model == "bank" else H.make_model(
        {"register": H.MODEL_REGISTER, "cas-register": H.MODEL_CAS_REGISTER}[model])
    seen = 0
    for seed in range(300):
        ops = arbitrary_history(model, np.random.default_rng(seed), max_events=16, n_proc=4)
        h = H.flatten_ops(ops, model)
        for eager in (False, True):
            r = oracle_mod.check_linearizable(h, m, 3, eager_reads=eager)
            f = oracle_mod.final_configs(h, m, 0, 1000, eager_reads=eager)
            if r["valid"] != H.INVALID:
                assert f["total"] == -1
                continue
            seen += 1
            wit_ok = r["shards"][0]["witness_index"]
            wit_inv = max(o["index"] for o in ops if o["type"] == "invoke" and o["index"] < wit_ok and
                          o["process"] == next(x["process"] for x in ops if x["index"] == wit_ok))
            assert f["total"] == len(f["configs"]) >= 1
            keys = [(c["state"], c["balances"], c["pending"], c["linearized_open"], c["crashed_linearized"])
                    for c in f["configs"]]
            for c in f["configs"]:
                assert wit_inv in c["pending"], (seed, c)
                assert not set(c["pending"]) & set(c["linearized_open"])
                if model == "bank":
                    assert sum(c["balances"]) == 0
            # canonical order: the struct's int32 fields in declaration order, unused entries 0
            flat = [[k[0]] + k[1] + [len(k[2]), len(k[3]), k[4]] + k[2] + [0] * (64 - len(k[2])) + k[3] +
                    [0] * (64 - len(k[3])) for k in keys]
            assert flat == sorted(flat)
    assert seen > 50
