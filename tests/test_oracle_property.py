"""Cross-checks between the four independent CPU deciders (CPU only).  brute == linear == wgl ==
wgl-compact on tiny histories; linear == wgl == wgl-compact on medium ones; metamorphic properties."""
import numpy as np
import pytest

from jepsen_tigerbeetle_b200 import history as H
from jepsen_tigerbeetle_b200 import synth

MODELS = {
    "register": lambda: H.make_model(H.MODEL_REGISTER),
    "cas-register": lambda: H.make_model(H.MODEL_CAS_REGISTER),
    "set": lambda: H.make_model(H.MODEL_SET),
    "bank": lambda: H.make_model(H.MODEL_BANK, accounts=range(1, 9)),
}


def verdict(r):
    return (r["valid"], r["shards"][0]["witness_index"])


@pytest.mark.parametrize("model", list(MODELS))
def test_tiny_all_four_agree(oracle_mod, model):
    n_invalid = 0
    for seed in range(60):
        spec = synth.SynthSpec(model, n_ops=7, n_clients=3, seed=seed, p_info=0.15 if seed % 2 else 0.0,
                               stale_read=seed % 3 != 0, stale_by=2 + seed % 3, n_values=3, n_accounts=3)
        h = synth.generate(spec)
        m = MODELS[model]() if model != "bank" else H.make_model(H.MODEL_BANK, accounts=range(1, 4))
        rs = [verdict(oracle_mod.check_linearizable(h, m, a)) for a in (0, 1, 2, 3, 4)]
        assert rs[0] == rs[1] == rs[2] == rs[3] == rs[4], (model, seed, rs)
        n_invalid += rs[0][0] == H.INVALID
    assert n_invalid > 0  # the mutation does produce invalid histories


@pytest.mark.parametrize("model", list(MODELS))
def test_medium_three_agree(oracle_mod, model):
    for seed in range(8):
        spec = synth.SynthSpec(model, n_ops=300, n_clients=5, seed=100 + seed, p_info=0.03 if seed % 2 else 0.0,
                               stale_read=seed % 2 == 0, n_values=4, tau_think_ns=5e6)
        h = synth.generate(spec)
        m = MODELS[model]()
        rs = [verdict(oracle_mod.check_linearizable(h, m, a)) for a in (1, 2, 3)]
        assert rs[0] == rs[1] == rs[2], (model, seed, rs)


def test_canonical_crashed_classes_do_not_change_verdict(oracle_mod):
    for seed in range(12):
        spec = synth.SynthSpec("cas-register", n_ops=120, n_clients=4, seed=500 + seed, p_info=0.3,
                               stale_read=seed % 2 == 0, n_values=3)
        h = synth.generate(spec)
        m = MODELS["cas-register"]()
        a = oracle_mod.check_linearizable(h, m, 3, canon_info=True)
        b = oracle_mod.check_linearizable(h, m, 3, canon_info=False)
        c = oracle_mod.check_linearizable(h, m, 2, canon_info=False)
        assert verdict(a) == verdict(b) == verdict(c)
        assert a["configs"] <= b["configs"]


def test_exhaustive_count_is_order_independent(oracle_mod):
    # for an INVALID history both WGL caches hold exactly the reachable configs
    spec = synth.SynthSpec("bank", n_ops=400, n_clients=6, seed=7, stale_read=True, tau_think_ns=2e6)
    h = synth.generate(spec)
    m = MODELS["bank"]()
    a = oracle_mod.check_linearizable(h, m, 2)
    b = oracle_mod.check_linearizable(h, m, 3)
    assert a["valid"] == b["valid"] == H.INVALID
    assert a["configs"] == b["configs"]


def test_metamorphic_process_relabel_and_time_dilation(oracle_mod):
    spec = synth.SynthSpec("cas-register", n_ops=200, n_clients=5, seed=11, stale_read=True)
    h = synth.generate(spec)
    m = MODELS["cas-register"]()
    base = verdict(oracle_mod.check_linearizable(h, m, 3))
    h2 = synth.generate(spec)
    h2.process[:] = (h2.process * 7 + 3).astype(np.int32)   # injective relabelling
    h2.time_ns[:] = h2.time_ns * 3 + 17
    assert verdict(oracle_mod.check_linearizable(h2, m, 3)) == base


def test_prefix_closure(oracle_mod):
    # a linearizable history stays linearizable when truncated (open invokes become crashed ops)
    spec = synth.SynthSpec("bank", n_ops=150, n_clients=4, seed=3)
    h = synth.generate(spec)
    m = MODELS["bank"]()
    assert oracle_mod.check_linearizable(h, m, 3)["valid"] == H.VALID
    for cut in (17, 60, 111, 250):
        hp = h.shard(0)
        sl = slice(0, cut)
        hp2 = H.FlatHistory(hp.type[sl].copy(), hp.f[sl].copy(), hp.flags[sl].copy(), hp.process[sl].copy(),
                            hp.index[sl].copy(), hp.time_ns[sl].copy(), hp.a[sl].copy(), hp.b[sl].copy(),
                            hp.c[sl].copy(), hp.payload_off[sl].copy(), hp.payload_len[sl].copy(), hp.payload,
                            np.array([0, cut], np.int64), hp.key_ids)
        assert oracle_mod.check_linearizable(hp2, m, 3)["valid"] == H.VALID


@pytest.mark.parametrize("model", list(MODELS))
def test_eager_reads_preserve_verdict_and_witness(oracle_mod, model):
    """The device's "eager reads" reduction (restated in the oracle as an option) never changes the answer."""
    for seed in range(40):
        spec = synth.SynthSpec(model, n_ops=9 if seed < 20 else 200, n_clients=3 if seed < 20 else 5, seed=seed,
                               p_info=0.15 if seed % 2 else 0.0, stale_read=seed % 3 != 0, stale_by=2 + seed % 3,
                               n_values=3, n_accounts=3 if seed < 20 else 8, tau_think_ns=0 if seed < 20 else 4e6)
        h = synth.generate(spec)
        m = MODELS[model]() if not (model == "bank" and seed < 20) else H.make_model(H.MODEL_BANK, accounts=range(1, 4))
        plain = oracle_mod.check_linearizable(h, m, 3)
        eager_c = oracle_mod.check_linearizable(h, m, 3, eager_reads=True)
        eager_f = oracle_mod.check_linearizable(h, m, 2, eager_reads=True)
        assert verdict(plain) == verdict(eager_c) == verdict(eager_f), (model, seed)
        assert eager_c["configs"] <= plain["configs"] or plain["valid"] != H.INVALID
        if seed < 20:
            assert verdict(oracle_mod.check_linearizable(h, m, 0)) == verdict(eager_c)  # vs brute force


@pytest.mark.parametrize("eager", [False, True])
@pytest.mark.parametrize("model", list(MODELS))
def test_level_decider_finds_every_duplicate_inside_its_level(oracle_mod, model, eager):
    """ALGO_LEVEL throws its visited set away after every level (the rule the device's level engine relies on: equal
    configurations always have equal depth).  On the oracle's own literal (BitSet, model) configurations it must reproduce
    knossos.wgl's verdict, witness and — for exhaustive searches — the exact number of distinct configurations."""
    n_counts = 0
    for seed in range(24):
        spec = synth.SynthSpec(model, n_ops=90, n_clients=5, seed=300 + seed, p_info=0.08 if seed % 2 else 0.0,
                               stale_read=seed % 3 != 0, stale_by=3 + seed % 4, n_values=3, tau_think_ns=3e6)
        h = synth.generate(spec)
        m = MODELS[model]()
        a = oracle_mod.check_linearizable(h, m, 3, eager_reads=eager)
        b = oracle_mod.check_linearizable(h, m, 4, eager_reads=eager)
        assert verdict(a) == verdict(b), (model, seed, eager)
        if a["valid"] == H.INVALID:
            assert a["configs"] == b["configs"], (model, seed, eager, a["configs"], b["configs"])
            n_counts += 1
    assert n_counts > 0


def test_lazy_bank_decider_equals_knossos_wgl(oracle_mod):
    """ALGO_LAZY_BANK (transfers linearized only when the frontier forces them or a read's balances require them) against
    knossos.wgl in the exact space — verdict, witness, previous-ok — and against brute force on tiny histories."""
    m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
    n_invalid = 0
    for seed in range(300):
        spec = synth.SynthSpec("bank", n_ops=30 + seed % 90, n_clients=2 + seed % 9, seed=seed, stale_read=seed % 2 == 0,
                               stale_by=2 + seed % 6, tau_think_ns=(seed % 5) * 2e6)
        h = synth.generate(spec)
        a, b = oracle_mod.check_linearizable(h, m, 3), oracle_mod.check_linearizable(h, m, 5)
        assert (a["valid"], a["shards"][0]["witness_index"], a["shards"][0]["previous_ok_index"]) == \
               (b["valid"], b["shards"][0]["witness_index"], b["shards"][0]["previous_ok_index"]), seed
        n_invalid += a["valid"] == H.INVALID
    assert n_invalid > 50
    m3 = H.make_model(H.MODEL_BANK, accounts=range(1, 4))
    for seed in range(80):
        spec = synth.SynthSpec("bank", n_ops=7, n_clients=3, seed=seed, stale_read=seed % 3 != 0, stale_by=2 + seed % 3, n_accounts=3)
        h = synth.generate(spec)
        assert verdict(oracle_mod.check_linearizable(h, m3, 0)) == verdict(oracle_mod.check_linearizable(h, m3, 5)), seed
    # what it is for: the bench headline (10k ops / 32 clients, tau_think 0), out of reach for every exhaustive CPU search
    h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=0, stale_read=True))
    r = oracle_mod.check_linearizable(h, m, 5)
    assert r["valid"] == H.INVALID and r["shards"][0]["witness_index"] == 18016 and r["configs"] < 1_000_000


def test_lazy_bank_decider_with_crashed_transfers(oracle_mod):
    """Crashed (:info) transfers enter the reduced search only as part of a read's set, class members in invocation
    order: verdict, witness and previous-ok must still equal knossos.wgl's."""
    m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
    n_invalid = 0
    for seed in range(160):
        spec = synth.SynthSpec("bank", n_ops=30 + seed % 70, n_clients=2 + seed % 8, seed=1000 + seed, stale_read=seed % 2 == 0,
                               stale_by=2 + seed % 6, tau_think_ns=(seed % 5) * 2e6, p_info=(0.05, 0.15, 0.3)[seed % 3])
        h = synth.generate(spec)
        a = oracle_mod.check_linearizable(h, m, 3, max_configs=3_000_000)
        if a["valid"] == H.UNKNOWN:
            continue
        b = oracle_mod.check_linearizable(h, m, 5, max_configs=3_000_000)
        assert (a["valid"], a["shards"][0]["witness_index"], a["shards"][0]["previous_ok_index"]) == \
               (b["valid"], b["shards"][0]["witness_index"], b["shards"][0]["previous_ok_index"]), seed
        n_invalid += a["valid"] == H.INVALID
    assert n_invalid > 20
