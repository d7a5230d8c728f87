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
