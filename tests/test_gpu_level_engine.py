"""The level-synchronous engine (csrc/jtb_level.cuh) forced on every kind of history — including those with crashed
(:info) ops and wide keys, which the default engine choice sends to the work list — against the CPU oracle: verdict,
witness, previous-ok and, for exhaustive single-key searches, the exact number of distinct configurations (the level
engine's visited set only lives for one level, so an equal count proves that no duplicate escapes it)."""
import numpy as np
import pytest

import arbitrary
import kat
from jepsen_tigerbeetle_b200 import history as H
from jepsen_tigerbeetle_b200 import synth

pytestmark = pytest.mark.gpu


def model_for(name, **kw):
    if name == "register":
        return H.make_model(H.MODEL_REGISTER)
    if name == "cas-register":
        return H.make_model(H.MODEL_CAS_REGISTER)
    if name == "set":
        return H.make_model(H.MODEL_SET)
    return H.make_model(H.MODEL_BANK, accounts=range(1, 9), **kw)


@pytest.fixture(scope="module")
def level_ctx():
    from jepsen_tigerbeetle_b200 import native
    with native.Context(device=0, engine="level") as ctx:
        yield ctx


@pytest.fixture(scope="module")
def level_ctx_exact():
    from jepsen_tigerbeetle_b200 import native
    with native.Context(device=0, engine="level", eager_reads=False) as ctx:
        yield ctx


def compare(g, o, counts=True):
    assert g["valid"] == o["valid"], (g, o)
    for gs, os_ in zip(g["shards"], o["shards"]):
        assert gs["valid"] == os_["valid"], (gs, os_)
        assert gs["witness_index"] == os_["witness_index"], (gs, os_)
        assert gs["previous_ok_index"] == os_["previous_ok_index"], (gs, os_)
    if counts and len(o["shards"]) == 1 and o["valid"] == H.INVALID:
        assert g["configs"] == o["configs"], (g["configs"], o["configs"])


def check(ctx, oracle_mod, h, m, eager, counts=True):
    g = ctx.check_linearizable(h, m)
    assert ctx.stats()["engine_level"] == 1
    compare(g, oracle_mod.check_linearizable(h, m, 3, eager_reads=eager), counts)
    return g


@pytest.mark.parametrize("name,model,text,expect,witness", kat.ALL_LIN_KATS, ids=[k[0] for k in kat.ALL_LIN_KATS])
def test_kats(level_ctx, level_ctx_exact, oracle_mod, name, model, text, expect, witness):
    h = H.flatten_ops(kat.ops(text), model)
    for ctx, eager in ((level_ctx, True), (level_ctx_exact, False)):
        g = check(ctx, oracle_mod, h, model_for(model), eager)
        assert g["valid"] == expect
        if expect == H.INVALID and witness is not None:
            assert g["shards"][0]["witness_index"] == witness


@pytest.mark.parametrize("model", ["register", "cas-register", "bank", "set"])
def test_random_small_with_crashed_ops(level_ctx, level_ctx_exact, oracle_mod, model):
    for seed in range(30):
        spec = synth.SynthSpec(model, n_ops=80, n_clients=5, seed=seed, p_info=0.1 if seed % 2 else 0.0,
                               stale_read=seed % 3 != 0, stale_by=3 + seed % 5, n_values=3)
        h = synth.generate(spec)
        check(level_ctx, oracle_mod, h, model_for(model), True)
        check(level_ctx_exact, oracle_mod, h, model_for(model), False)


@pytest.mark.parametrize("model", ["register", "cas-register", "set", "bank"])
def test_arbitrary_small_histories(level_ctx, oracle_mod, model):
    rng = np.random.default_rng(4242)
    m = H.make_model(H.MODEL_BANK, accounts=[1, 2, 3]) if model == "bank" else model_for(model)
    for _ in range(150):
        ops = arbitrary.arbitrary_history(model, rng)
        h = H.flatten_ops(ops, model)
        check(level_ctx, oracle_mod, h, m, True)


@pytest.mark.parametrize("p_info", [0.0, 0.05])
@pytest.mark.parametrize("stale", [False, True])
def test_config_c2(level_ctx, level_ctx_exact, oracle_mod, p_info, stale):
    h = synth.config_c2(seed=1, p_info=p_info, stale_read=stale)
    m = model_for("cas-register")
    if p_info > 0 and oracle_mod.check_linearizable(h, m, 3, eager_reads=True)["valid"] == H.VALID:
        # a VALID history with crashed ops: breadth-first visits the whole reachable space (3 x 10^10 configurations
        # here, where the depth-first work list needs 10^4) - which is why the default engine choice sends histories
        # with crashed ops to the work list.  Forced, the level engine must give up cleanly, never contradict.
        from jepsen_tigerbeetle_b200 import native
        with native.Context(device=0, engine="level", max_configs=20_000_000) as ctx:
            g = ctx.check_linearizable(h, m)
        assert g["valid"] in (H.VALID, H.UNKNOWN)
        return
    check(level_ctx, oracle_mod, h, m, True)
    if p_info == 0.0:
        check(level_ctx_exact, oracle_mod, h, m, False)


@pytest.mark.parametrize("stale", [False, True])
def test_bank_3000_ops_wide_levels(level_ctx, level_ctx_exact, oracle_mod, stale):
    """bank 3000 ops / 32 clients (the golden case): 5.8 M configurations, levels up to 577 k wide -> grid barriers,
    windows of several MB, staging flushes; exact count in both spaces."""
    h = synth.generate(synth.SynthSpec("bank", 3000, 32, 2, tau_think_ns=20e6, stale_read=stale))
    m = model_for("bank")
    g = check(level_ctx, oracle_mod, h, m, True)
    ge = check(level_ctx_exact, oracle_mod, h, m, False)
    if stale:
        assert ge["configs"] > 20 * g["configs"]
        assert level_ctx_exact.stats()["ring_tail"] > 100_000     # (stat slot 3 = widest level in level mode)


def test_multi_key_with_a_poisoned_key(level_ctx, oracle_mod):
    h = synth.poison_c5(synth.config_c5(seed=3, n_keys=24, n_ops=4000, p_info=0.05), 7)
    m = model_for("cas-register")
    g = check(level_ctx, oracle_mod, h, m, True)
    assert g["valid"] == H.INVALID and g["n_failures"] == 1


@pytest.mark.parametrize("model,n_clients,n_ops,think", [("cas-register", 48, 600, 4e6), ("bank", 40, 500, 6e6)])
def test_more_than_32_open_ops(level_ctx, oracle_mod, model, n_clients, n_ops, think):
    for stale in (False, True):
        h = synth.generate(synth.SynthSpec(model, n_ops, n_clients, 5, tau_think_ns=think, stale_read=stale, n_values=40))
        check(level_ctx, oracle_mod, h, model_for(model), True)


@pytest.mark.parametrize("n_clients,n_ops,p_info,n_values", [(12, 900, 0.4, 14), (8, 1500, 0.4, 24)])
def test_wide_keys_many_crashed_op_classes(level_ctx, oracle_mod, n_clients, n_ops, p_info, n_values):
    """32 / 64 B keys (lock-bit protocol inside the level window) on an INVALID history: exhaustive count."""
    h = synth.generate(synth.SynthSpec("cas-register", n_ops, n_clients, 3, p_info=p_info, n_values=n_values,
                                       stale_read=True, tau_think_ns=30e6))
    m = model_for("cas-register")
    o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True, max_configs=30_000_000)
    if o["valid"] == H.UNKNOWN:
        pytest.skip("oracle budget")
    g = level_ctx.check_linearizable(h, m)
    assert g["key_bytes"] >= 32
    compare(g, o)


def test_growth_of_window_and_level_arrays(oracle_mod, monkeypatch):
    """Start with a 1 MiB window and 1 MiB level arrays: the kernel stops at the level that does not fit, the host
    grows both 4x, carries the unfinished level over and relaunches — the exhaustive count must not change."""
    from jepsen_tigerbeetle_b200 import native
    monkeypatch.setenv("JTB_LV_TABLE_MB", "1")
    monkeypatch.setenv("JTB_LV_BUF_MB", "1")
    h = synth.generate(synth.SynthSpec("bank", 3000, 32, 2, tau_think_ns=20e6, stale_read=True))
    m = model_for("bank")
    with native.Context(device=0, engine="level", eager_reads=False) as ctx:
        g = ctx.check_linearizable(h, m)
        assert ctx.stats()["attempts"] > 1
    compare(g, oracle_mod.check_linearizable(h, m, 3))


def test_budgets_give_unknown(oracle_mod):
    from jepsen_tigerbeetle_b200 import native
    h = synth.generate(synth.SynthSpec("bank", 3000, 32, 2, tau_think_ns=20e6, stale_read=True))
    m = model_for("bank")
    with native.Context(device=0, engine="level", eager_reads=False, max_configs=100_000) as ctx:
        g = ctx.check_linearizable(h, m)
        assert g["valid"] == H.UNKNOWN and g["shards"][0]["cause"] == 2
    with native.Context(device=0, engine="level", eager_reads=False, time_budget_ms=1) as ctx:
        g = ctx.check_linearizable(h, m)
        assert g["valid"] in (H.UNKNOWN, H.INVALID)


def test_final_configs_after_a_level_search(oracle_mod):
    from jepsen_tigerbeetle_b200 import native
    h = synth.generate(synth.SynthSpec("bank", 600, 8, 2, tau_think_ns=5e6, stale_read=True))
    m = model_for("bank")
    with native.Context(device=0, engine="level") as ctx:
        g = ctx.check_linearizable(h, m)
        assert g["valid"] == H.INVALID
        fc = ctx.final_configs(h, m, 0, cap=10)
    o = oracle_mod.final_configs(h, m, 0, cap=10, eager_reads=True)
    assert fc["total"] == o["total"] and fc["configs"] == o["configs"]


CRASH_HEAVY_VALID = [
    synth.SynthSpec('cas-register', 2500, 24, 809007372, p_info=0.3, tau_think_ns=20e6, n_values=30),
    synth.SynthSpec('register', 1000, 24, 902980068, p_info=0.3, tau_think_ns=5e6, n_values=30, stale_read=True),
    synth.SynthSpec('register', 2500, 40, 321354213, p_info=0.1, tau_think_ns=20e6, n_values=30, stale_by=3),
    synth.SynthSpec('cas-register', 1000, 16, 1, p_info=0.05),
    synth.SynthSpec('cas-register', 50000, 64, 1, p_info=0.3, n_keys=8, grouped_keys=True),      # C5 "monster": 8 keys: no beam
]


@pytest.mark.parametrize("spec", CRASH_HEAVY_VALID, ids=[f"{s.model}-{s.n_ops}-{s.n_clients}-{s.p_info}" for s in CRASH_HEAVY_VALID])
def test_beam_finds_the_linearization_of_crash_heavy_valid_histories(oracle_mod, spec):
    """VALID histories with 5-30 % crashed ops: round 1 answered :unknown (breadth-first crowd) or needed 0.6-12 s of
    depth-first scouts.  The beam (level engine keeping the best ~1k / ~16k configurations of every level) decides them."""
    from jepsen_tigerbeetle_b200 import native
    h = synth.generate(spec)
    m = model_for(spec.model)
    assert oracle_mod.check_linearizable(h, m, 3, eager_reads=True, max_configs=5_000_000, n_threads=8)["valid"] == H.VALID
    with native.Context(device=0, time_budget_ms=60_000) as ctx:
        g = ctx.check_linearizable(h, m)
        st = ctx.stats()
    assert g["valid"] == H.VALID, (g, st)
    if spec.model == "register":          # beyond the 16 M-configuration work-list probe: decided by the beam alone
        assert st["beam_decided"] == 1 and st["scouts"] == 0, st


def test_beam_never_decides_an_invalid_history(oracle_mod):
    """A beam can only report VALID: an INVALID history with crashed ops falls through to the exhaustive search, whose
    verdict, witness and configuration count are those of the oracle."""
    from jepsen_tigerbeetle_b200 import native
    h = synth.generate(synth.SynthSpec("bank", 3000, 16, 1, p_info=0.02, stale_read=True, tau_think_ns=4e6))
    m = model_for("bank")
    o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True, max_configs=50_000_000)
    assert o["valid"] == H.INVALID
    with native.Context(device=0) as ctx:
        g = ctx.check_linearizable(h, m)
        st = ctx.stats()
    assert st["beam_decided"] == 0, st
    compare(g, o)
