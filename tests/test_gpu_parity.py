"""GPU parity tests proper: every call goes through the C ABI (libjtb_check.so) and is compared with
the CPU oracle on the same seeded inputs — bit-exact verdict, witness index and (for exhaustive
searches) the number of distinct configurations."""
import numpy as np
import pytest

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


def same_verdict(g, o):
    assert g["valid"] == o["valid"], (g, o)
    for gs, os_ in zip(g["shards"], o["shards"]):
        assert gs["valid"] == os_["valid"]
        assert gs["witness_index"] == os_["witness_index"], (gs, os_)
        assert gs["previous_ok_index"] == os_["previous_ok_index"], (gs, os_)


GPU_LIN_KATS = list(kat.ALL_LIN_KATS)


@pytest.mark.parametrize("name,model,text,expect,witness", GPU_LIN_KATS, ids=[k[0] for k in GPU_LIN_KATS])
def test_kats(gpu_ctx, oracle_mod, name, model, text, expect, witness):
    h = H.flatten_ops(kat.ops(text), model)
    g = gpu_ctx.check_linearizable(h, model_for(model))
    assert g["valid"] == expect
    if expect == H.INVALID and witness is not None:
        assert g["shards"][0]["witness_index"] == witness
    same_verdict(g, oracle_mod.check_linearizable(h, model_for(model), 3, eager_reads=True))


def test_bank_negative_balances_forbidden(gpu_ctx):
    h = H.flatten_ops(kat.ops("0:inv transfer t(1 2 3), 0:ok transfer t(1 2 3)"), "bank")
    assert gpu_ctx.check_linearizable(h, model_for("bank", negative_balances_ok=False))["valid"] == H.INVALID
    assert gpu_ctx.check_linearizable(h, model_for("bank"))["valid"] == H.VALID


@pytest.mark.parametrize("model", ["register", "cas-register", "bank", "set"])
def test_random_small(gpu_ctx, oracle_mod, model):
    for seed in range(40):
        spec = synth.SynthSpec(model, n_ops=60, n_clients=4, seed=seed, p_info=0.1 if seed % 2 else 0.0,
                               stale_read=seed % 3 != 0, stale_by=3 + seed % 5, n_values=3)
        h = synth.generate(spec)
        m = model_for(model)
        g = gpu_ctx.check_linearizable(h, m)
        o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True)
        same_verdict(g, o)
        if o["valid"] == H.INVALID:
            assert g["configs"] == o["configs"], (model, seed)


@pytest.mark.parametrize("p_info", [0.0, 0.05])
@pytest.mark.parametrize("stale", [False, True])
def test_config_c2(gpu_ctx, oracle_mod, p_info, stale):
    """BASELINE config #2: 1k-op cas-register history, 16 clients."""
    for seed in (1, 2, 3):
        h = synth.config_c2(seed=seed, p_info=p_info, stale_read=stale)
        m = model_for("cas-register")
        g = gpu_ctx.check_linearizable(h, m)
        o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True)
        same_verdict(g, o)
        if o["valid"] == H.INVALID:
            assert g["configs"] == o["configs"]


@pytest.mark.parametrize("stale", [False, True])
def test_config_c3_lite(gpu_ctx, oracle_mod, stale):
    """BASELINE config #3 shape at a size the oracle finishes in seconds: bank, 32 clients."""
    h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=20e6, stale_read=stale))
    m = model_for("bank")
    g = gpu_ctx.check_linearizable(h, m)
    o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True)
    same_verdict(g, o)
    if stale:
        assert o["valid"] == H.INVALID and g["configs"] == o["configs"]


def test_bank_with_crashed_transfers(gpu_ctx, oracle_mod):
    for seed in (1, 2):
        h = synth.generate(synth.SynthSpec("bank", 600, 8, seed, p_info=0.05, tau_think_ns=10e6,
                                           stale_read=seed == 2))
        m = model_for("bank")
        g = gpu_ctx.check_linearizable(h, m)
        o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True)
        same_verdict(g, o)


def test_multi_shard(gpu_ctx, oracle_mod):
    """independent keys: per-key verdicts, merged with merge-valid; one poisoned key flips the verdict."""
    h = synth.generate(synth.SynthSpec("cas-register", 4000, 64, 5, p_info=0.1, n_keys=8, grouped_keys=True))
    m = model_for("cas-register")
    g = gpu_ctx.check_linearizable(h, m)
    o = oracle_mod.check_linearizable(h, m, 3, n_threads=4, eager_reads=True)
    same_verdict(g, o)
    assert g["valid"] == H.VALID
    h = synth.generate(synth.SynthSpec("cas-register", 4000, 64, 5, p_info=0.1, n_keys=8, grouped_keys=True,
                                       stale_read=True))
    g = gpu_ctx.check_linearizable(h, m)
    o = oracle_mod.check_linearizable(h, m, 3, n_threads=4, eager_reads=True)
    same_verdict(g, o)
    assert g["n_failures"] == o["n_failures"]


def test_set_model_keyed_c4_lite(gpu_ctx, oracle_mod):
    """BASELINE config #4 shape (set-full workload, many ledgers) under the knossos set model."""
    for seed, stale in ((1, False), (2, True)):
        h = synth.generate(synth.SynthSpec("set", 6000, 32, seed, p_info=0.01, n_keys=8, stale_read=stale,
                                           tau_think_ns=5e6))
        m = model_for("set")
        g = gpu_ctx.check_linearizable(h, m)
        o = oracle_mod.check_linearizable(h, m, 3, n_threads=4, eager_reads=True)
        same_verdict(g, o)
        if stale:
            assert o["valid"] == H.INVALID


def test_budget_gives_unknown(oracle_mod):
    from jepsen_tigerbeetle_b200 import native
    h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=20e6, stale_read=True))
    with native.Context(max_configs=20000) as ctx:
        g = ctx.check_linearizable(h, model_for("bank"))
    assert g["valid"] == H.UNKNOWN and g["shards"][0]["cause"] == 2


# ---- scans -------------------------------------------------------------------------------------------
def sf_equal(g, o):
    assert g["valid"] == o["valid"]
    assert g["shards"] == o["shards"]
    assert g["raia_valid"] == o["raia_valid"] and g["suspect_final_reads"] == o["suspect_final_reads"]
    for k in ("elem_off", "elem_id", "elem_outcome", "elem_latency_ms", "elem_dup_count"):
        assert np.array_equal(g[k], o[k]), k


def test_set_full_c1(gpu_ctx, oracle_mod):
    for seed in (1, 2, 3):
        h = synth.config_c1(seed=seed)
        for lin in (True, False):
            sf_equal(gpu_ctx.check_set_full(h, lin), oracle_mod.check_set_full(h, lin))


SF_KATS = [
    ("stable", "0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{1}", [0, 1, 2, 3]),
    ("never-read", "0:inv add 1, 0:ok add 1", [0, 1]),
    ("lost", "0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{1}, 1:inv read, 1:ok read #{}", [0, 1, 2, 3, 4, 5]),
    ("stale", "0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{}, 1:inv read, 1:ok read #{1}", [0, 1, 2, 3, 12, 13]),
    ("read-before-add-ok", "0:inv add 1, 1:inv read, 1:ok read #{1}, 0:ok add 1", [0, 1, 2, 3]),
    ("info-add-seen", "0:inv add 1, 0:info add 1, 1:inv read, 1:ok read #{1}", [0, 1, 2, 3]),
    ("tracked-duplicate", "0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read [1 1]", [0, 1, 2, 3]),
    # an id that was never :add-invoked occurs twice in one read: jepsen counts (frequencies v) over every value
    ("untracked-duplicate", "0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read [1 7 7]", [0, 1, 2, 3]),
    ("untracked-single", "0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{1 7}", [0, 1, 2, 3]),
    # sparse ids: the id -> position lookup falls back from the direct table to the sorted table
    ("sparse-ids", "0:inv add 5, 0:ok add 5, 0:inv add 900000, 0:ok add 900000, 0:inv add 70000000, 0:ok add 70000000, "
                   "1:inv read, 1:ok read #{5 70000000}, 1:inv read, 1:ok read #{5 900000 70000000}", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]),
    ("re-added-element", "0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{1}, 0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{}",
     [0, 1, 2, 3, 4, 5, 6, 7]),
]


@pytest.mark.parametrize("name,text,times", SF_KATS, ids=[k[0] for k in SF_KATS])
def test_set_full_kats(gpu_ctx, oracle_mod, name, text, times):
    h = H.flatten_ops(kat.ops(text, [t * 1_000_000 for t in times]), "set")
    for lin in (True, False):
        sf_equal(gpu_ctx.check_set_full(h, lin), oracle_mod.check_set_full(h, lin))


def test_set_full_pinned_payload_and_buffer_reuse(gpu_ctx, oracle_mod):
    """Page-locked id lists (jtb_host_alloc) and the context's cached device buffers: a large history, then a small
    one, then the large one again through the same context."""
    from jepsen_tigerbeetle_b200 import native
    big = native.pin_history(synth.config_c4(seed=3, n_keys=4, n_ops=12000))
    small = synth.config_c1(seed=2)
    ob, os_ = oracle_mod.check_set_full(big), oracle_mod.check_set_full(small)
    for h, o in ((big, ob), (small, os_), (big, ob)):
        sf_equal(gpu_ctx.check_set_full(h), o)


def test_set_full_multi_key_with_info(gpu_ctx, oracle_mod):
    h = synth.config_c4(seed=2, n_keys=8, n_ops=8000)
    sf_equal(gpu_ctx.check_set_full(h), oracle_mod.check_set_full(h))


def test_bank_totals(gpu_ctx, oracle_mod):
    h = synth.generate(synth.SynthSpec("bank", 3000, 16, 4, tau_think_ns=20e6))
    m = model_for("bank")
    g, o = gpu_ctx.check_bank_totals(h, m, 0), oracle_mod.check_bank_totals(h, m, 0)
    for k in g:
        if not k.startswith("seconds"):
            assert g[k] == o[k], k
    # corrupt some reads
    rng = np.random.default_rng(0)
    idx = np.flatnonzero(h.payload_len > 0)
    for e in rng.choice(idx, 40, replace=False):
        h.payload[h.payload_off[e] + 1 + 2 * int(rng.integers(0, 8))] += int(rng.integers(-9, 9))
    for neg_ok in (True, False):
        m = model_for("bank", negative_balances_ok=neg_ok)
        g, o = gpu_ctx.check_bank_totals(h, m, 0), oracle_mod.check_bank_totals(h, m, 0)
        for k in g:
            if not k.startswith("seconds"):
                assert g[k] == o[k], (k, neg_ok)


def test_bank_totals_total_amount_zero_quirk(gpu_ctx, oracle_mod):
    """tests/ledger.clj:122-123 with :total-amount 0: two :wrong-total reads make the reference throw -> :unknown."""
    from jepsen_tigerbeetle_b200 import checker as ck
    z = "3 0 4 0 5 0 6 0 7 0 8 0"
    one = f"0:inv read, 0:ok read {{1 1 2 0 {z}}}"
    two = one + f", 0:inv read, 0:ok read {{1 -7 2 0 {z}}}"
    m = model_for("bank")
    for text, total, expect in ((one, 0, (H.INVALID, 0)), (two, 0, (H.UNKNOWN, 1)), (two, 10, (H.INVALID, 0))):
        h = H.flatten_ops(kat.ops(text), "bank")
        g, o = gpu_ctx.check_bank_totals(h, m, total), oracle_mod.check_bank_totals(h, m, total)
        assert (g["valid"], g["reference_throws"]) == expect
        for k in g:
            if not k.startswith("seconds"):
                assert g[k] == o[k], k
    r = ck.bank_checker({"negative-balances?": True}, ctx=gpu_ctx).check(
        {"accounts": list(range(1, 9)), "total-amount": 0}, kat.ops(two))
    assert r["valid?"] == "unknown" and "Divide by zero" in r["error"] and r["error-count"] == 2


def test_read_all_invoked_adds(gpu_ctx, oracle_mod):
    """workloads/set_full.clj:51-75 on the device: final reads that miss invoked adds."""
    h = synth.config_c4(seed=3, n_keys=4, n_ops=4000)
    sf_equal(gpu_ctx.check_set_full(h), oracle_mod.check_set_full(h))
    finals = np.flatnonzero(((h.flags & 1) != 0) & (h.type == 1))
    before = {s["index"]: len(s["missing"]) for s in gpu_ctx.check_set_full(h)["suspect_final_reads"]}
    for e in finals[:2]:
        h.payload_len[e] -= 3          # the final read of two ledgers loses its last three elements
    g, o = gpu_ctx.check_set_full(h), oracle_mod.check_set_full(h)
    sf_equal(g, o)
    after = {s["index"]: len(s["missing"]) for s in g["suspect_final_reads"]}
    assert g["raia_valid"] == H.INVALID
    # (never-applied crashed adds are legitimately missing already; the truncation adds exactly three)
    assert sorted(after[i] - before.get(i, 0) for i in after) == [0, 0, 3, 3]


def test_checker_protocol_end_to_end(gpu_ctx):
    """The reference's compose maps, through the Python mirror of the Checker protocol."""
    from jepsen_tigerbeetle_b200 import checker as ck
    h = synth.config_c4(seed=4, n_keys=3, n_ops=1500, p_info=0.0)
    c = ck.independent_checker(ck.compose({"set-full": ck.set_full({"linearizable?": True}, ctx=gpu_ctx),
                                           "read-all-invoked-adds": ck.read_all_invoked_adds(ctx=gpu_ctx),
                                           "linear": ck.linearizable({"model": "set"}, ctx=gpu_ctx)}))
    r = ck.check_safe(c, {}, h)
    assert set(r["results"]) == {1, 2, 3}
    for k, m in r["results"].items():
        assert set(m) == {"set-full", "read-all-invoked-adds", "linear", "valid?"}
        assert m["linear"]["valid?"] is True and m["read-all-invoked-adds"]["valid?"] is True
    hb = synth.generate(synth.SynthSpec("bank", 800, 8, 2, tau_think_ns=10e6, stale_read=True))
    cb = ck.compose({"SI": ck.bank_checker({"negative-balances?": True}, ctx=gpu_ctx),
                     "linear": ck.linearizable({"model": "bank"}, ctx=gpu_ctx)})
    rb = ck.check_safe(cb, {"accounts": list(range(1, 9)), "total-amount": 0}, hb)
    assert rb["SI"]["valid?"] is True          # totals are preserved by a stale read ...
    assert rb["linear"]["valid?"] is False     # ... but the history is not linearizable
    assert rb["valid?"] is False and rb["linear"]["op"]["index"] >= 0


def test_knossos_exact_mode_without_eager_reads(oracle_mod):
    """JTB_OPT_NO_EAGER_READS: the device visits exactly the configurations knossos.wgl's search space holds."""
    from jepsen_tigerbeetle_b200 import native
    with native.Context(eager_reads=False) as ctx:
        for model, spec in (("bank", synth.SynthSpec("bank", 3000, 16, 3, tau_think_ns=10e6, stale_read=True)),
                            ("cas-register", synth.SynthSpec("cas-register", 1000, 16, 1, n_values=30, stale_read=True)),
                            ("set", synth.SynthSpec("set", 800, 8, 2, stale_read=True, tau_think_ns=5e6))):
            h = synth.generate(spec)
            m = model_for(model)
            g = ctx.check_linearizable(h, m)
            o = oracle_mod.check_linearizable(h, m, 3)          # plain WGL, no reduction
            same_verdict(g, o)
            assert o["valid"] == H.INVALID and g["configs"] == o["configs"], model
            e = oracle_mod.check_linearizable(h, m, 3, eager_reads=True)
            assert e["configs"] <= o["configs"]


@pytest.mark.parametrize("model,n_clients,n_ops,think", [("cas-register", 48, 600, 4e6), ("bank", 40, 500, 6e6),
                                                         ("set", 44, 500, 5e6)])
def test_more_than_32_open_ops(gpu_ctx, oracle_mod, model, n_clients, n_ops, think):
    """> 32 concurrently open ops: 64 slot lanes, two candidate rounds per expansion."""
    from jepsen_tigerbeetle_b200 import native
    for stale in (False, True):
        h = synth.generate(synth.SynthSpec(model, n_ops, n_clients, 1 + stale, tau_think_ns=think, stale_read=stale,
                                           n_values=6))
        m = model_for(model)
        assert native.prepare_info(h, m)["slot_lanes"] == 64
        g = gpu_ctx.check_linearizable(h, m)
        o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True)
        same_verdict(g, o)
        if o["valid"] == H.INVALID:
            assert g["configs"] == o["configs"]


@pytest.mark.parametrize("n_clients,n_ops,p_info,n_values,key_bytes", [(12, 900, 0.4, 14, 32), (8, 1500, 0.4, 24, 64)])
def test_wide_keys_many_crashed_op_classes(gpu_ctx, oracle_mod, n_clients, n_ops, p_info, n_values, key_bytes):
    """Hundreds of crashed-op classes: 32 B / 64 B keys (lock-bit slots), several class rounds per expansion."""
    from jepsen_tigerbeetle_b200 import native
    m = model_for("cas-register")
    h = synth.generate(synth.SynthSpec("cas-register", n_ops, n_clients, 3, tau_think_ns=2e6, p_info=p_info,
                                       n_values=n_values))
    assert native.prepare_info(h, m)["key_bytes"] == key_bytes
    g = gpu_ctx.check_linearizable(h, m)
    assert g["key_bytes"] == key_bytes
    same_verdict(g, oracle_mod.check_linearizable(h, m, 3, eager_reads=True))
    assert g["valid"] == H.VALID
    # an :ok read of a value nobody ever wrote, early in the history: invalid, exhaustive up to that read
    reads = np.flatnonzero((h.type == 1) & (h.f == 0) & (h.a != H.NIL))
    h.a[reads[3]] = 999
    g = gpu_ctx.check_linearizable(h, m)
    o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True)
    same_verdict(g, o)
    assert g["valid"] == H.INVALID and g["shards"][0]["witness_index"] == int(h.index[reads[3]])
    assert g["configs"] == o["configs"]


def test_time_budget_and_edge_histories(gpu_ctx):
    from jepsen_tigerbeetle_b200 import native
    m = model_for("bank")
    h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=2e6, stale_read=True))
    with native.Context(time_budget_ms=5, eager_reads=False) as ctx:      # 5 ms is far too little for this search
        g = ctx.check_linearizable(h, m)
    assert g["valid"] == H.UNKNOWN and g["shards"][0]["cause"] == 2 and g["shards"][0]["witness_index"] == -1
    # a history with no completed op at all (every op crashed) is trivially linearizable
    ops = kat.ops("0:inv write 1, 1:inv write 2, 0:info write 1")
    g = gpu_ctx.check_linearizable(H.flatten_ops(ops, "cas-register"), model_for("cas-register"))
    assert g["valid"] == H.VALID and g["configs"] == 0
    # only nemesis ops
    g = gpu_ctx.check_linearizable(H.flatten_ops([{"type": "info", "f": "kill", "process": "nemesis", "value": None}],
                                                 "cas-register"), model_for("cas-register"))
    assert g["valid"] == H.VALID
    # malformed: completion without invocation -> error, which the Checker mirror turns into :unknown
    bad = H.flatten_ops(kat.ops("0:ok write 1"), "cas-register")
    with pytest.raises(native.NativeError):
        gpu_ctx.check_linearizable(bad, model_for("cas-register"))
    from jepsen_tigerbeetle_b200 import checker as ck
    r = ck.check_safe(ck.linearizable({"model": "cas-register"}, ctx=gpu_ctx), {}, bad)
    assert r["valid?"] == "unknown" and "completion without invocation" in r["error"]


@pytest.mark.parametrize("model", ["register", "cas-register", "set", "bank"])
def test_arbitrary_small_histories(gpu_ctx, oracle_mod, model):
    """300 arbitrary (mostly non-linearizable) small histories per model: verdict, witness and — when the search
    is exhaustive — the configuration count must equal the oracle's (which agrees with brute force on these)."""
    import arbitrary
    rng = np.random.default_rng(20260922)
    m = H.make_model(H.MODEL_BANK, accounts=[1, 2, 3]) if model == "bank" else model_for(model)
    n_invalid = 0
    for _ in range(300):
        ops = arbitrary.arbitrary_history(model, rng)
        h = H.flatten_ops(ops, model)
        g = gpu_ctx.check_linearizable(h, m)
        o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True)
        same_verdict(g, o)
        if o["valid"] == H.INVALID:
            n_invalid += 1
            assert g["configs"] == o["configs"], ops
            assert oracle_mod.check_linearizable(h, m, 0)["shards"][0]["witness_index"] == g["shards"][0]["witness_index"]
    assert n_invalid > 30


def test_pause_resume_on_ring_and_table_growth(oracle_mod, monkeypatch):
    """Pause/resume: the live work is exactly the non-zero ring slots.  A deliberately tiny ring guard forces the
    RING_FULL pause -> flush -> compact into a 4x ring -> relaunch path; a 32 MiB table forces several re-hashes.
    The exhaustive configuration count must still equal the oracle's (no work lost or duplicated)."""
    from jepsen_tigerbeetle_b200 import native
    h = synth.generate(synth.SynthSpec("bank", 10000, 32, 1, tau_think_ns=10e6, stale_read=True))
    m = model_for("bank")
    monkeypatch.setenv("JTB_TEST_TINY_RING", "1")
    monkeypatch.setenv("JTB_TABLE_START_MB", "32")
    for eager in (True, False):
        o = oracle_mod.check_linearizable(h, m, 3, eager_reads=eager)
        assert o["valid"] == H.INVALID
        with native.Context(eager_reads=eager, engine="worklist") as ctx:
            g = ctx.check_linearizable(h, m)
            st = ctx.stats()
        same_verdict(g, o)
        assert g["configs"] == o["configs"] and g["probes"] == o["probes"], st
        assert st["attempts"] >= 2, st
        if not eager:  # the wide exact-space search overruns the tiny ring guard as well
            assert st["attempts"] >= 4 and st["ring_entries"] > (1 << 22), st


SCOUT_CASES = [("register", dict(n_ops=400, n_clients=6, p_info=0.1, n_values=5)),
               ("cas-register", dict(n_ops=600, n_clients=8, p_info=0.2, n_values=5, tau_think_ns=2e6)),
               ("cas-register", dict(n_ops=2500, n_clients=24, p_info=0.3, n_values=30, tau_think_ns=20e6)),
               ("bank", dict(n_ops=300, n_clients=6, p_info=0.1, tau_think_ns=4e6)),
               ("set", dict(n_ops=300, n_clients=5, p_info=0.1)),
               ("register", dict(n_ops=3000, n_clients=4, p_info=0.0, n_keys=8))]


@pytest.mark.gpu
@pytest.mark.parametrize("model,kw", SCOUT_CASES, ids=[f"{m}-{k['n_ops']}" for m, k in SCOUT_CASES])
def test_scout_walks_the_cpu_depth_first_order(oracle_mod, monkeypatch, model, kw):
    """A depth-first scout in order 0 is knossos.wgl's walk: run ALONE (no search kernel) on a valid history it must
    find the linearization after inserting exactly the configs the CPU port inserts (the port also counts the final,
    complete config, one per shard; the scout reports success before inserting it)."""
    from jepsen_tigerbeetle_b200 import native
    h = synth.generate(synth.SynthSpec(model, seed=11, **kw))
    m = model_for(model)
    monkeypatch.setenv("JTB_SCOUT_ONLY", "1")
    monkeypatch.setenv("JTB_SCOUTS", "1")
    monkeypatch.setenv("JTB_SCOUT_ORDERS", "1")
    for eager in (True, False):
        o = oracle_mod.check_linearizable(h, m, 3, eager_reads=eager)
        assert o["valid"] == H.VALID
        with native.Context(eager_reads=eager) as ctx:
            g = ctx.check_linearizable(h, m)
            st = ctx.stats()
        assert g["valid"] == H.VALID, (g, st)
        assert st["scouts"] == 1 and st["scout_decided"] == h.n_shards, st
        assert st["scout_configs"] == o["configs"] - h.n_shards, (st, o["configs"])


@pytest.mark.gpu
def test_scouts_rescue_crash_heavy_valid_histories(oracle_mod):
    """Found by the soak test: valid histories with 10-30 % crashed ops, where breadth-first exploration runs out of
    budget (:unknown) but a depth-first walk finds the linearization.  With scouts the verdict is VALID; without,
    the same budget gives UNKNOWN (never a wrong answer)."""
    from jepsen_tigerbeetle_b200 import native
    specs = [synth.SynthSpec('cas-register', 2500, 24, 809007372, p_info=0.3, tau_think_ns=20e6, n_values=30),
             synth.SynthSpec('register', 1000, 24, 902980068, p_info=0.3, tau_think_ns=5e6, n_values=30, stale_read=True)]
    for sp in specs:
        h = synth.generate(sp)
        m = model_for(sp.model)
        assert oracle_mod.check_linearizable(h, m, 3, eager_reads=True, max_configs=5_000_000)["valid"] == H.VALID
        with native.Context(max_configs=50_000_000, beam=False) as ctx:
            g = ctx.check_linearizable(h, m)
            st = ctx.stats()
        assert g["valid"] == H.VALID, (g, st)
        assert st["scouts"] == 4, st
        with native.Context(max_configs=50_000_000, scouts=False, beam=False) as ctx:
            g0 = ctx.check_linearizable(h, m)
            assert ctx.stats()["scouts"] == 0
        assert g0["valid"] in (H.VALID, H.UNKNOWN)


@pytest.mark.gpu
def test_scouts_do_not_disturb_exhaustive_counts_or_growth(oracle_mod, monkeypatch):
    """An INVALID history with crashed ops: scouts run beside the search (they can never decide it), the table is
    re-hashed several times while they run (deferred frees), and the exhaustive count still equals the oracle's."""
    from jepsen_tigerbeetle_b200 import native
    h = synth.generate(synth.SynthSpec("bank", 3000, 16, 1, p_info=0.02, stale_read=True, tau_think_ns=4e6))
    m = model_for("bank")
    o = oracle_mod.check_linearizable(h, m, 3, eager_reads=True, max_configs=50_000_000)
    assert o["valid"] == H.INVALID
    monkeypatch.setenv("JTB_TABLE_START_MB", "1")
    with native.Context(beam=False) as ctx:      # (the default would settle this one in its budgeted work-list probe)
        g = ctx.check_linearizable(h, m)
        st = ctx.stats()
    same_verdict(g, o)
    assert g["configs"] == o["configs"], (st, o["configs"])
    assert st["scouts"] == 4 and st["scout_decided"] == 0, st


FINAL_CONFIG_CASES = [
    ("register", (1, 4), dict(n_ops=400, n_clients=6, p_info=0.02, n_values=30, stale_read=True, stale_by=3)),
    ("cas-register", (1,), dict(n_ops=1000, n_clients=16, n_values=30, stale_read=True)),
    ("bank", (1, 3), dict(n_ops=1500, n_clients=10, p_info=0.03, stale_read=True, tau_think_ns=4e6)),
    ("register", (2, 3), dict(n_ops=3000, n_clients=4, n_values=30, stale_read=True, n_keys=8)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("model,seeds,kw", FINAL_CONFIG_CASES, ids=[f"{m}-{k['n_ops']}" for m, _s, k in FINAL_CONFIG_CASES])
def test_final_configs_match_the_oracle(oracle_mod, model, seeds, kw):
    """knossos' :configs: every visited configuration stuck at the witness, read back from the
    visited table in HBM and decoded, equals the oracle's list config for config (state, balances, pending ops,
    linearized-but-open ops, crashed ops consumed), in both search modes, for single- and multi-key histories."""
    from jepsen_tigerbeetle_b200 import native
    m = model_for(model)
    checked = 0
    for seed in seeds:
        h = synth.generate(synth.SynthSpec(model, seed=seed, **kw))
        for eager in (True, False):
            o = oracle_mod.check_linearizable(h, m, 3, eager_reads=eager, max_configs=30_000_000)
            assert o["valid"] == H.INVALID
            with native.Context(eager_reads=eager) as ctx:
                g = ctx.check_linearizable(h, m)
                same_verdict(g, o)
                for k, s in enumerate(o["shards"]):
                    if s["valid"] != H.INVALID:
                        with pytest.raises(native.NativeError):
                            ctx.final_configs(h, m, shard=k)
                        continue
                    fg = ctx.final_configs(h, m, shard=k, cap=20000)
                    fo = oracle_mod.final_configs(h, m, shard=k, cap=100000, eager_reads=eager)
                    assert fg["total"] == fo["total"] >= 1, (seed, eager, k)
                    assert fg["configs"] == fo["configs"], (seed, eager, k)
                    assert ctx.final_configs(h, m, shard=k, cap=3)["configs"] == fo["configs"][:3]
                    checked += 1
                if model == "bank":   # any other call on the context invalidates the table of the last search
                    ctx.check_bank_totals(h, m)
                    with pytest.raises(native.NativeError):
                        ctx.final_configs(h, m, shard=0)
    assert checked >= 2


@pytest.mark.gpu
def test_linearizable_checker_reports_configs(gpu_ctx):
    """The Checker-protocol mirror returns knossos' :configs for an invalid history (first 10, like jepsen)."""
    from jepsen_tigerbeetle_b200 import checker

    def op(p, t, f, v, i):
        return {"process": p, "type": t, "f": f, "value": v, "index": i, "time": i * 1000}
    hist = [op(0, "invoke", "write", 1, 0), op(0, "ok", "write", 1, 1), op(1, "invoke", "write", 3, 2),
            op(0, "invoke", "read", None, 3), op(0, "ok", "read", 2, 4), op(1, "ok", "write", 3, 5)]
    r = checker.linearizable({"model": "register"}).check({}, hist, {})
    assert r["valid?"] is False and r["op"] == {"index": 4} and r["configs-total"] == 2
    assert r["configs"] == [
        {"model": 1, "pending": [{"index": 2}, {"index": 3}], "linearized-open": [], "crashed-linearized": 0},
        {"model": 3, "pending": [{"index": 3}], "linearized-open": [{"index": 2}], "crashed-linearized": 0}]


def test_device_partition_by_key_and_ledger_balances(gpu_ctx):
    """independent/subhistory as ONE stable device partition (vs numpy's stable argsort) and ledger->bank's
    balance arithmetic (tests/ledger.clj:100-105)."""
    rng = np.random.default_rng(7)
    for n, nk in ((0, 1), (1, 1), (1000, 7), (200_000, 64), (300_000, 5000)):
        keys = rng.choice(np.concatenate([rng.integers(-5, 50, nk), rng.integers(-2**62, 2**62, 3)]), size=n).astype(np.int64) if n else np.zeros(0, np.int64)
        r = gpu_ctx.partition_by_key(keys)
        order = np.argsort(keys, kind="stable").astype(np.int32)
        assert np.array_equal(r["order"], order)
        ids, first = np.unique(keys[order], return_index=True) if n else (np.zeros(0, np.int64), np.zeros(0, np.int64))
        assert np.array_equal(r["key_ids"], ids)
        assert np.array_equal(r["shard_off"], np.concatenate([first, [n]]).astype(np.int64))
    c = rng.integers(0, 10**9, 100_000); d = rng.integers(0, 10**9, 100_000)
    assert np.array_equal(gpu_ctx.ledger_balances(c, d), (c - d).astype(np.int32))
