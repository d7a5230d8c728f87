"""CPU tier: host logic of the checker-protocol mirror (compose / independent / check-safe / merge-valid)
with stub inner checkers — no GPU call is made here."""
import pytest

from jepsen_tigerbeetle_b200 import checker as ck
from jepsen_tigerbeetle_b200 import history as H
import kat


class Const(ck.Checker):
    def __init__(self, v):
        self.v = v

    def check(self, test, history, opts=None):
        if isinstance(self.v, Exception):
            raise self.v
        return {"valid?": self.v, "key": (opts or {}).get("history-key")}


def test_merge_valid_lattice():
    assert ck.merge_valid_values([True, True]) is True
    assert ck.merge_valid_values([True, "unknown"]) == "unknown"
    assert ck.merge_valid_values([True, "unknown", False]) is False
    assert ck.merge_valid_values([]) is True


def test_compose_merges_and_check_safe_degrades_to_unknown():
    c = ck.compose({"a": Const(True), "b": Const(RuntimeError("boom"))})
    r = c.check({}, [])
    assert r["a"]["valid?"] is True and r["b"]["valid?"] == "unknown" and "boom" in r["b"]["error"]
    assert r["valid?"] == "unknown"
    assert ck.compose({"a": Const(True), "b": Const(False)}).check({}, [])["valid?"] is False


def test_independent_splits_by_key_and_reports_failures():
    hist = []
    for k, text in ((1, "0:inv add 1, 0:ok add 1"), (7, "1:inv add 5, 1:ok add 5")):
        for op in kat.ops(text):
            op = dict(op)
            op["value"] = (k, op["value"])
            hist.append(op)
    seen = []

    class PerKey(ck.Checker):
        def check(self, test, history, opts=None):
            seen.append((opts["history-key"], history.n_events))
            return {"valid?": opts["history-key"] != 7}

    r = ck.independent_checker(PerKey(), model="set").check({}, hist)
    assert sorted(seen) == [(1, 2), (7, 2)]
    assert r["valid?"] is False and r["failures"] == [7] and set(r["results"]) == {1, 7}


def test_linearizable_requires_a_model():
    with pytest.raises(AssertionError):
        ck.Linearizable("no-such-model")


def test_flattener_accepts_both_bank_spellings():
    a = H.flatten_ops([{"type": "invoke", "f": "transfer", "process": 0, "value": {"from": 1, "to": 2, "amount": 3}}], "bank")
    b = H.flatten_ops([{"type": "invoke", "f": "transfer", "process": 0,
                        "value": {"debit-acct": 1, "credit-acct": 2, "amount": 3}}], "bank")
    assert (a.a[0], a.b[0], a.c[0]) == (b.a[0], b.b[0], b.c[0]) == (3, 1, 2)


# ---- the ledger test's host-side checkers (tests/ledger.clj:194-282) --------------------------------------
def _t(i, p, typ, micro, final=False, t=None):
    op = {"index": i, "process": p, "type": typ, "f": "txn", "value": micro, "time": t if t is not None else i * 1000}
    if final:
        op["final?"] = True
    return op


def test_unexpected_ops():
    tr = [["t", 1, {"debit-acct": 1, "credit-acct": 2, "amount": 3}]]
    h = [_t(0, 0, "invoke", tr), _t(1, 0, "ok", tr), {"index": 2, "process": "nemesis", "type": "info", "f": "kill", "value": None, "time": 2000}]
    assert ck.unexpected_ops().check({}, h) == {"valid?": True}
    h2 = h + [_t(3, 1, "invoke", tr, t=5_000_000), _t(4, 2, "invoke", tr, t=6_000_000), _t(5, 2, "fail", tr, t=9_000_000)]
    r = ck.unexpected_ops().check({}, h2)
    assert r["valid?"] == "unknown" and len(r["fail-ops"]) == 1
    assert [o["index"] for _, o in r["open-ops"]] == [3] and r["open-ops"][0][0] == 4.0  # (9e6 - 5e6) ns -> ms


def test_lookup_all_invoked_transfers_and_final_reads():
    t1 = [["t", 1, {"debit-acct": 1, "credit-acct": 2, "amount": 3}]]
    t2 = [["t", 2, {"debit-acct": 2, "credit-acct": 3, "amount": 1}]]
    rd = [["r", a, {"credits-posted": 0, "debits-posted": 0}] for a in (1, 2)]
    h = [_t(0, 0, "invoke", t1), _t(1, 0, "ok", t1), _t(2, 1, "invoke", t2), _t(3, 1, "info", t2),
         _t(4, 0, "invoke", [["r", 1, None], ["r", 2, None]], final=True), _t(5, 0, "ok", rd, final=True),
         _t(6, 1, "invoke", [["r", 1, None], ["r", 2, None]], final=True), _t(7, 1, "ok", rd, final=True),
         _t(8, 0, "invoke", [["l-t", None, None]], final=True), _t(9, 0, "ok", [["l-t", 1, {}], ["l-t", 2, {}]], final=True)]
    assert ck.lookup_all_invoked_transfers().check({}, h) == {"valid?": True}
    assert ck.final_reads().check({}, h) == {"valid?": True}
    h_bad = h[:-1] + [_t(9, 0, "ok", [["l-t", 1, {}]], final=True)]          # the crashed transfer 2 is missing
    r = ck.lookup_all_invoked_transfers().check({}, h_bad)
    assert r["valid?"] is False and [o["index"] for o in r["suspect-final-lookups"]] == [9]
    rd2 = [["r", 1, {"credits-posted": 1, "debits-posted": 0}], ["r", 2, {"credits-posted": 0, "debits-posted": 0}]]
    h_uneq = h[:7] + [_t(7, 1, "ok", rd2, final=True)] + h[8:]
    r = ck.final_reads().check({}, h_uneq)
    assert r["valid?"] is False and len(r["unequal-final-reads"]) == 2
    r = ck.final_reads().check({}, h[:4])                                       # no final reads at all
    assert r["valid?"] is False and r["unequal-final-reads"] == set() and r["unequal-final-lookups"] == set()


def test_stats_checker():
    """(checker/stats), core.clj:144: counts by completion type, overall and by :f; invalid when some :f never succeeded."""
    from jepsen_tigerbeetle_b200 import checker
    hist = [
        {"type": "invoke", "f": "add", "process": 0, "value": 1}, {"type": "ok", "f": "add", "process": 0, "value": 1},
        {"type": "invoke", "f": "add", "process": 1, "value": 2}, {"type": "info", "f": "add", "process": 1, "value": 2},
        {"type": "invoke", "f": "read", "process": 0}, {"type": "fail", "f": "read", "process": 0},
        {"type": "info", "f": "start", "process": "nemesis"},
    ]
    r = checker.stats().check({}, hist)
    assert (r["count"], r["ok-count"], r["fail-count"], r["info-count"]) == (3, 1, 1, 1)
    assert r["by-f"]["add"] == {"valid?": True, "count": 2, "ok-count": 1, "fail-count": 0, "info-count": 1}
    assert r["by-f"]["read"]["valid?"] is False and r["valid?"] is False
    hist.append({"type": "invoke", "f": "read", "process": 0})
    hist.append({"type": "ok", "f": "read", "process": 0, "value": [1]})
    assert checker.stats().check({}, hist)["valid?"] is True
