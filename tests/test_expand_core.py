"""CPU tier: the DEVICE expansion core (csrc/jtb_expand.h, compiled for the host by tests/native/hostwalk.cpp) against
the oracle — verdict, witness and, for exhaustive searches of single-key histories, the configuration count."""
import numpy as np
import pytest

import arbitrary
import hostwalk
import kat
from jepsen_tigerbeetle_b200 import history as H
from jepsen_tigerbeetle_b200 import synth


def model_for(name, **kw):
    if name == "register":
        return H.make_model(H.MODEL_REGISTER)
    if name == "cas-register":
        return H.make_model(H.MODEL_CAS_REGISTER)
    if name == "set":
        return H.make_model(H.MODEL_SET)
    return H.make_model(H.MODEL_BANK, accounts=range(1, 9), **kw)


def compare(w, o, counts=True):
    assert w["valid"] == o["valid"], (w, o)
    for ws, os_ in zip(w["shards"], o["shards"]):
        if os_["cause"] == 3:      # too wide for the device encoding: the walker reports UNKNOWN as well
            assert ws["valid"] == H.UNKNOWN
            continue
        assert ws["valid"] == os_["valid"], (ws, os_)
        assert ws["witness_index"] == os_["witness_index"], (ws, os_)
        assert ws["previous_ok_index"] == os_["previous_ok_index"], (ws, os_)
    if counts and len(o["shards"]) == 1 and o["valid"] == H.INVALID:
        assert w["configs"] == o["configs"], (w["configs"], o["configs"])


@pytest.mark.parametrize("name,model,text,expect,witness", kat.ALL_LIN_KATS, ids=[k[0] for k in kat.ALL_LIN_KATS])
def test_kats(oracle_mod, name, model, text, expect, witness):
    h = H.flatten_ops(kat.ops(text), model)
    for eager in (True, False):
        w = hostwalk.walk(h, model_for(model), eager_reads=eager)
        assert w["valid"] == expect
        if expect == H.INVALID and witness is not None:
            assert w["shards"][0]["witness_index"] == witness
        compare(w, oracle_mod.check_linearizable(h, model_for(model), 3, eager_reads=eager))


@pytest.mark.parametrize("eager", [True, False])
@pytest.mark.parametrize("model", ["register", "cas-register", "bank", "set"])
def test_random_small(oracle_mod, model, eager):
    for seed in range(40):
        spec = synth.SynthSpec(model, n_ops=60, n_clients=4, seed=seed, p_info=0.1 if seed % 2 else 0.0,
                               stale_read=seed % 3 != 0, stale_by=3 + seed % 5, n_values=3)
        h = synth.generate(spec)
        m = model_for(model)
        compare(hostwalk.walk(h, m, eager_reads=eager), oracle_mod.check_linearizable(h, m, 3, eager_reads=eager))


@pytest.mark.parametrize("model", ["register", "cas-register", "bank", "set"])
def test_arbitrary_small_histories(oracle_mod, model):
    rng = np.random.default_rng(20260923)
    for _ in range(300):
        ops = arbitrary.arbitrary_history(model, rng)
        h = H.flatten_ops(ops, model)
        m = H.make_model(H.MODEL_BANK, accounts=[1, 2, 3]) if model == "bank" else model_for(model)
        for eager in (True, False):
            compare(hostwalk.walk(h, m, eager_reads=eager), oracle_mod.check_linearizable(h, m, 3, eager_reads=eager))


def test_bank_negative_balances_forbidden():
    h = H.flatten_ops(kat.ops("0:inv transfer t(1 2 3), 0:ok transfer t(1 2 3)"), "bank")
    assert hostwalk.walk(h, model_for("bank", negative_balances_ok=False))["valid"] == H.INVALID
    assert hostwalk.walk(h, model_for("bank"))["valid"] == H.VALID


@pytest.mark.parametrize("stale", [False, True])
def test_bank_32_clients(oracle_mod, stale):
    h = synth.generate(synth.SynthSpec("bank", 3000, 32, 1, tau_think_ns=20e6, stale_read=stale))
    m = model_for("bank")
    for eager in (True, False):
        compare(hostwalk.walk(h, m, eager_reads=eager), oracle_mod.check_linearizable(h, m, 3, eager_reads=eager))


def test_crashed_ops_wide_keys_and_64_slots(oracle_mod):
    # crashed-op classes push the key to 32 / 64 bytes; 48 clients need the 64-slot rows
    for model, n_ops, clients, p_info, seed in (("cas-register", 300, 6, 0.3, 3), ("register", 400, 48, 0.02, 2),
                                                ("bank", 300, 6, 0.2, 4), ("set", 400, 8, 0.1, 5)):
        for stale in (False, True):
            h = synth.generate(synth.SynthSpec(model, n_ops, clients, seed, p_info=p_info, stale_read=stale,
                                               tau_think_ns=10e6))
            m = model_for(model)
            o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True, max_configs=3_000_000)
            if o["valid"] == H.UNKNOWN:
                continue
            compare(hostwalk.walk(h, m, eager_reads=True, max_configs=30_000_000), o)


def test_multi_key(oracle_mod):
    for stale in (False, True):
        h = synth.generate(synth.SynthSpec("cas-register", 2000, 32, 5, p_info=0.05, n_keys=8, grouped_keys=True,
                                           stale_read=stale))
        m = model_for("cas-register")
        compare(hostwalk.walk(h, m), oracle_mod.check_linearizable(h, m, 3, n_threads=4, eager_reads=True), counts=False)


# ---- the level engine's rule on the CPU: breadth-first by depth, visited set local to a level, candidate reads decided
#      from the rows' summary words (csrc/jtb_level.cuh runs exactly this expansion core) --------------------------------
@pytest.mark.parametrize("name,model,text,expect,witness", kat.ALL_LIN_KATS, ids=[k[0] for k in kat.ALL_LIN_KATS])
def test_level_walk_kats(oracle_mod, name, model, text, expect, witness):
    h = H.flatten_ops(kat.ops(text), model)
    for eager in (True, False):
        w = hostwalk.walk_bfs(h, model_for(model), eager_reads=eager)
        assert w["valid"] == expect
        compare(w, oracle_mod.check_linearizable(h, model_for(model), 3, eager_reads=eager))


@pytest.mark.parametrize("eager", [True, False])
@pytest.mark.parametrize("model", ["register", "cas-register", "bank", "set"])
def test_level_walk_random_small(oracle_mod, model, eager):
    for seed in range(40):
        spec = synth.SynthSpec(model, n_ops=60, n_clients=4, seed=seed, p_info=0.1 if seed % 2 else 0.0,
                               stale_read=seed % 3 != 0, stale_by=3 + seed % 5, n_values=3)
        h = synth.generate(spec)
        m = model_for(model)
        compare(hostwalk.walk_bfs(h, m, eager_reads=eager), oracle_mod.check_linearizable(h, m, 3, eager_reads=eager))


def test_level_walk_counts_on_a_wide_bank_history(oracle_mod):
    """5.8 M configurations, levels up to 577 k wide: a per-level visited set finds every duplicate."""
    h = synth.generate(synth.SynthSpec("bank", 3000, 32, 2, tau_think_ns=20e6, stale_read=True))
    m = model_for("bank")
    w = hostwalk.walk_bfs(h, m, eager_reads=True, width_cap=4000)
    o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True)
    compare(w, o)
    assert w["levels"] > 2000 and max(w["widths"]) > 5000
