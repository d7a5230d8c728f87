"""The oracle against the hand known-answer tests of SURVEY.md Appendix B (CPU only).

The reference ships no golden vectors for this path (test/tigerbeetle/core_test.clj:4-6 asserts
nothing), so these first-principles KATs + the cross-implementation property tests are what pins
the oracle ("parity unpinned" — see oracle/oracle_common.h)."""
import numpy as np
import pytest

import kat
from jepsen_tigerbeetle_b200 import history as H

ALGOS = [0, 1, 2, 3, 4]  # brute, linear, wgl (full bitset), wgl compact, level (per-level visited set)


def model_for(name, **kw):
    if name == "register":
        return H.make_model(H.MODEL_REGISTER)
    if name == "cas-register":
        return H.make_model(H.MODEL_CAS_REGISTER)
    if name == "set":
        return H.make_model(H.MODEL_SET)
    return H.make_model(H.MODEL_BANK, accounts=range(1, 9), **kw)


@pytest.mark.parametrize("name,model,text,expect,witness", kat.ALL_LIN_KATS, ids=[k[0] for k in kat.ALL_LIN_KATS])
@pytest.mark.parametrize("algo", ALGOS)
def test_linearizable_kat(oracle_mod, name, model, text, expect, witness, algo):
    h = H.flatten_ops(kat.ops(text), model)
    r = oracle_mod.check_linearizable(h, model_for(model), algo)
    assert r["valid"] == expect, (name, r)
    if expect == H.INVALID and witness is not None:
        assert r["shards"][0]["witness_index"] == witness, (name, r)
    if expect == H.VALID:
        assert r["shards"][0]["witness_index"] == -1


@pytest.mark.parametrize("algo", ALGOS)
def test_bank_negative_balances_forbidden(oracle_mod, algo):
    # B44: with negative balances forbidden a transfer from a zero balance cannot happen
    text = "0:inv transfer t(1 2 3), 0:ok transfer t(1 2 3)"
    h = H.flatten_ops(kat.ops(text), "bank")
    r = oracle_mod.check_linearizable(h, model_for("bank", negative_balances_ok=False), algo)
    assert r["valid"] == H.INVALID and r["shards"][0]["witness_index"] == 1
    r = oracle_mod.check_linearizable(h, model_for("bank", negative_balances_ok=True), algo)
    assert r["valid"] == H.VALID


def test_ledger_form_maps_to_bank(oracle_mod):
    # B45 / ledger->bank (tests/ledger.clj:89-114): balance = credits-posted - debits-posted
    hist = [
        {"type": "invoke", "f": "txn", "process": 0, "value": [["t", 1, {"debit-acct": 1, "credit-acct": 2, "amount": 3}]]},
        {"type": "ok", "f": "txn", "process": 0, "value": [["t", 1, {"debit-acct": 1, "credit-acct": 2, "amount": 3}]]},
        {"type": "invoke", "f": "txn", "process": 1, "value": [["r", a, None] for a in range(1, 9)]},
        {"type": "ok", "f": "txn", "process": 1, "value": [["r", 1, {"credits-posted": 5, "debits-posted": 8}],
                                                          ["r", 2, {"credits-posted": 3, "debits-posted": 0}]] +
                                                         [["r", a, {"credits-posted": 0, "debits-posted": 0}] for a in range(3, 9)]},
        {"type": "invoke", "f": "txn", "process": 2, "value": [["l-t", None, None]]},
        {"type": "ok", "f": "txn", "process": 2, "value": [["l-t", 1, {}]]},
        {"type": "info", "f": "start", "process": "nemesis", "value": None},
    ]
    h = H.flatten_ops(hist, "bank")
    assert h.n_events == 4  # l-t and nemesis ops dropped
    assert list(h.payload[:4]) == [1, -3, 2, 3]
    r = oracle_mod.check_linearizable(h, model_for("bank"), 3)
    assert r["valid"] == H.VALID
    t = oracle_mod.check_bank_totals(h, model_for("bank"), 0)
    assert t["valid"] == H.VALID and t["read_count"] == 1


# ---- set-full; times in ns ------------------------------------------
MS = 1_000_000


def sf(text, times):
    return H.flatten_ops(kat.ops(text, times), "set")


def test_set_full_kats(oracle_mod):
    o = oracle_mod
    # B30 stable, not stale
    r = o.check_set_full(sf("0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{1}", [0, 1 * MS, 2 * MS, 3 * MS]))
    assert r["valid"] == H.VALID and r["shards"][0]["stable_count"] == 1 and r["shards"][0]["stale_count"] == 0
    # B31 never read -> :unknown
    r = o.check_set_full(sf("0:inv add 1, 0:ok add 1", [0, 1 * MS]))
    assert r["valid"] == H.UNKNOWN and r["shards"][0]["never_read_count"] == 1
    # B32 lost
    r = o.check_set_full(sf("0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{1}, 1:inv read, 1:ok read #{}",
                            [0, MS, 2 * MS, 3 * MS, 4 * MS, 5 * MS]))
    assert r["valid"] == H.INVALID and r["shards"][0]["lost_count"] == 1
    assert list(r["elem_outcome"]) == [2]
    # B33 stale by 5 ms: invalid only under :linearizable? true
    text = "0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{}, 1:inv read, 1:ok read #{1}"
    t33 = [-MS, 0, 5 * MS, 6 * MS, 10 * MS, 11 * MS]
    r = o.check_set_full(sf(text, t33), True)
    assert r["valid"] == H.INVALID and r["shards"][0]["stale_count"] == 1 and list(r["elem_latency_ms"]) == [5]
    assert o.check_set_full(sf(text, t33), False)["valid"] == H.VALID
    # B34 staleness of 0.4 ms truncates to 0 ms -> not stale (recall-sensitive KAT)
    t34 = [-MS, 0, 400_000, 500_000, 10 * MS, 11 * MS]
    r = o.check_set_full(sf(text, t34), True)
    assert r["valid"] == H.VALID and r["shards"][0]["stale_count"] == 0
    # B35 known via the read's ok when it precedes the add's ok
    r = o.check_set_full(sf("0:inv add 1, 1:inv read, 1:ok read #{1}, 0:ok add 1", [0, MS, 2 * MS, 3 * MS]))
    assert r["valid"] == H.VALID and r["shards"][0]["stable_count"] == 1
    # B36 crashed add observed by a read
    r = o.check_set_full(sf("0:inv add 1, 0:info add 1, 1:inv read, 1:ok read #{1}", [0, MS, 2 * MS, 3 * MS]))
    assert r["valid"] == H.VALID
    # B37 duplicate
    r = o.check_set_full(sf("0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read [1 1]", [0, MS, 2 * MS, 3 * MS]))
    assert r["valid"] == H.INVALID and r["shards"][0]["duplicated_count"] == 1 and list(r["elem_dup_count"]) == [2]
    # element read but never added is ignored (B24 gap)
    r = o.check_set_full(sf("0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{1 7}", [0, MS, 2 * MS, 3 * MS]))
    assert r["valid"] == H.VALID and r["shards"][0]["attempt_count"] == 1


def test_set_full_independent_keys(oracle_mod):
    # B38: key 1 = B30 (valid), key 2 = B32 (lost) -> merged false, failures = [key 2]
    hist = []
    for k, text in ((1, "0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{1}"),
                    (2, "2:inv add 5, 2:ok add 5, 3:inv read, 3:ok read #{5}, 3:inv read, 3:ok read #{}")):
        for op in kat.ops(text):
            op = dict(op)
            op["value"] = (k, op["value"])
            hist.append(op)
    for i, op in enumerate(hist):
        op["index"] = i
        op["time"] = i * MS
    h = H.flatten_ops(hist, "set")
    assert h.n_shards == 2 and list(h.key_ids) == [1, 2]
    r = oracle_mod.check_set_full(h)
    assert [s["valid"] for s in r["shards"]] == [H.VALID, H.INVALID]
    assert r["valid"] == H.INVALID and r["n_failures"] == 1


# ---- bank SI totals (tests/ledger.clj:127-192) -----------------------------------------------------
def test_bank_totals_kats(oracle_mod):
    m = model_for("bank")
    z = "3 0 4 0 5 0 6 0 7 0 8 0"
    def run(text, **kw):
        return oracle_mod.check_bank_totals(H.flatten_ops(kat.ops(text), "bank"), model_for("bank", **kw), 0)
    assert run(f"0:inv read, 0:ok read {{1 -3 2 3 {z}}}")["valid"] == H.VALID            # B40
    r = run(f"0:inv read, 0:ok read {{1 -3 2 0 {z}}}")                                       # B41
    assert r["valid"] == H.INVALID and r["first_error_type"] == 3 and r["lowest_total"] == -3
    r = run(f"0:inv read, 0:ok read {{1 0 2 0 {z} 9 0}}")                                    # B43
    assert r["first_error_type"] == 1
    r = run(f"0:inv read, 0:ok read {{1 -3 2 3 {z}}}", negative_balances_ok=False)           # B44
    assert r["first_error_type"] == 4
    r = run(f"0:inv read, 0:ok read {{1 nil 2 0 {z}}}")
    assert r["first_error_type"] == 2
    # precedence: unexpected-key beats nil-balance beats wrong-total (cond order, ledger.clj:132-152)
    r = run(f"0:inv read, 0:ok read {{1 nil 2 5 {z} 9 0}}")
    assert r["first_error_type"] == 1 and r["error_count"] == 1
    # aggregation: counts, first/last, lowest/highest
    r = run(f"0:inv read, 0:ok read {{1 1 2 0 {z}}}, 0:inv read, 0:ok read {{1 -7 2 0 {z}}}, 0:inv read, 0:ok read {{1 0 2 0 {z}}}, "
            f"0:inv read, 0:ok read {{1 4 2 0 {z}}}")
    assert r["read_count"] == 4 and r["error_count"] == 3 and r["count_by_type"][3] == 3
    assert r["first_index_by_type"][3] == 1 and r["last_index_by_type"][3] == 7
    assert (r["lowest_total"], r["lowest_index"], r["highest_total"], r["highest_index"]) == (-7, 3, 4, 7)
    assert r["worst_index_by_type"][3] == 3 and r["first_error_index"] == 1
    del m


def test_bank_totals_total_amount_zero_quirk(oracle_mod):
    """tests/ledger.clj:122-123: err-badness computes (/ (- total total-amount) total-amount); with the default
    :total-amount 0 (tests/ledger.clj:356) that is an integer division by zero.  util/max-by only calls it when an
    error type has >= 2 members, so ONE :wrong-total read gives {:valid? false} but TWO make the reference checker
    throw, which jepsen's check-safe reports as {:valid? :unknown}.  Oracle and device reproduce exactly that."""
    z = "3 0 4 0 5 0 6 0 7 0 8 0"
    def run(text, total=0):
        return oracle_mod.check_bank_totals(H.flatten_ops(kat.ops(text), "bank"), model_for("bank"), total)
    one = f"0:inv read, 0:ok read {{1 1 2 0 {z}}}"
    two = one + f", 0:inv read, 0:ok read {{1 -7 2 0 {z}}}"
    r = run(one)
    assert (r["valid"], r["reference_throws"], r["error_count"]) == (H.INVALID, 0, 1)
    r = run(two)
    assert (r["valid"], r["reference_throws"], r["error_count"]) == (H.UNKNOWN, 1, 2)
    assert r["count_by_type"][3] == 2 and r["worst_index_by_type"][3] == 3     # statistics are still filled
    r = run(two, total=10)                                                      # a non-zero total: no division by zero
    assert (r["valid"], r["reference_throws"]) == (H.INVALID, 0)
    # other error types never divide: two :nil-balance reads stay {:valid? false}
    r = run(f"0:inv read, 0:ok read {{1 nil 2 0 {z}}}, 0:inv read, 0:ok read {{1 nil 2 nil {z}}}")
    assert (r["valid"], r["reference_throws"], r["count_by_type"][2]) == (H.INVALID, 0, 2)
