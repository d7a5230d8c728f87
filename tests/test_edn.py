"""CPU tier: the EDN reader on op shapes the reference writes, through flatten + oracle."""
from jepsen_tigerbeetle_b200 import edn
from jepsen_tigerbeetle_b200 import history as H

SET_FULL_EDN = """
; set_full.clj:29-31,114-134 shapes
{:type :invoke, :f :add, :value [1 9], :time 3291485317, :process 0, :node "n1", :client [0 0], :index 0}
{:type :ok, :f :add, :value [1 9], :time 3296209422, :process 0, :index 1}
{:type :info, :f :start, :value nil, :time 3400000000, :process :nemesis, :index 2}
{:type :invoke, :f :read, :value [1 nil], :time 3500000000, :process 1, :index 3}
{:type :ok, :f :read, :value [1 #{9}], :time 3600000000, :process 1, :index 4}
{:type :invoke, :f :read, :value [1 nil], :final? true, :time 9500000000, :process 1, :index 5}
{:type :info, :f :read, :value [1 nil], :error :timeout, :final? true, :time 9600000000, :process 1, :index 6}
"""

LEDGER_EDN = """
[{:type :invoke, :f :txn, :value [[:t 1 {:debit-acct 1, :credit-acct 2, :amount 3N}]], :process 0, :time 10, :index 0}
 {:type :ok, :f :txn, :value [[:t 1 {:debit-acct 1, :credit-acct 2, :amount 3}]], :process 0, :time 20, :index 1}
 #jepsen.history.Op{:type :invoke, :f :txn, :value [[:r 1 nil] [:r 2 nil]], :process 1, :time 30, :index 2}
 {:type :ok, :f :txn, :value [[:r 1 {:credits-posted 0, :debits-posted 3}] [:r 2 {:credits-posted 3, :debits-posted 0}]], :process 1, :time 40, :index 3}
 {:type :invoke, :f :txn, :value [[:l-t nil nil]], :process 2, :time 50, :index 4} #_ {:ignored :form}
 {:type :ok, :f :txn, :value [[:l-t 1 {}]], :process 2, :time 60, :index 5}]
"""


def test_atoms_and_collections():
    assert edn.loads('{:a 1, :b [1 2.5 -3], :c #{:x "y\\"z"}, :d nil, :e true}') == {
        "a": 1, "b": [1, 2.5, -3], "c": frozenset({"x", 'y"z'}), "d": None, "e": True}
    assert edn.loads("#inst \"2022-10-01T00:00:00Z\"") == "2022-10-01T00:00:00Z"
    assert list(edn.loads_all("1 :k ; comment\n [2]")) == [1, "k", [2]]


def test_set_full_history_roundtrip(oracle_mod):
    ops = edn.read_history(SET_FULL_EDN, independent=True)
    assert len(ops) == 7 and ops[0]["value"] == (1, 9) and ops[2]["process"] == "nemesis"
    h = H.flatten_ops(ops, "set")
    assert h.n_events == 6 and h.n_shards == 1 and list(h.key_ids) == [1]   # nemesis op dropped
    r = oracle_mod.check_set_full(h, True)
    assert r["valid"] == H.VALID and r["shards"][0]["stable_count"] == 1
    assert oracle_mod.check_linearizable(h, H.make_model(H.MODEL_SET), 3)["valid"] == H.VALID


def test_ledger_history_roundtrip(oracle_mod):
    ops = edn.read_history(LEDGER_EDN)
    assert len(ops) == 6
    h = H.flatten_ops(ops, "bank")
    assert h.n_events == 4                      # :l-t ops dropped by ledger->bank
    assert list(h.payload) == [1, -3, 2, 3]     # balance = credits-posted - debits-posted
    m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
    assert oracle_mod.check_linearizable(h, m, 3)["valid"] == H.VALID
    assert oracle_mod.check_bank_totals(h, m, 0)["valid"] == H.VALID


# ---- the same stored-history shapes through the GPU (history.edn -> op maps -> flat arrays -> C ABI) ----
import pytest  # noqa: E402


def _edn_of(ops):
    """Op maps -> the EDN text jepsen's store writes (one map per line)."""
    def val(v):
        if v is None:
            return "nil"
        if isinstance(v, bool):
            return "true" if v else "false"
        if isinstance(v, str):
            return ":" + v
        if isinstance(v, (list, tuple)):
            return "[" + " ".join(val(x) for x in v) + "]"
        if isinstance(v, (set, frozenset)):
            return "#{" + " ".join(val(x) for x in sorted(v)) + "}"
        if isinstance(v, dict):
            return "{" + ", ".join(f":{k} {val(x)}" for k, x in v.items()) + "}"
        return str(v)
    return "\n".join("{" + ", ".join(f":{k} {val(v)}" for k, v in op.items()) + "}" for op in ops)


@pytest.mark.gpu
def test_stored_edn_histories_through_the_gpu(gpu_ctx, oracle_mod):
    ops = edn.read_history(SET_FULL_EDN, independent=True)
    h = H.flatten_ops(ops, "set")
    g, o = gpu_ctx.check_set_full(h, True), oracle_mod.check_set_full(h, True)
    assert g["valid"] == o["valid"] == H.VALID and g["shards"] == o["shards"]
    assert gpu_ctx.check_linearizable(h, H.make_model(H.MODEL_SET))["valid"] == H.VALID
    ops = edn.read_history(LEDGER_EDN)
    h = H.flatten_ops(ops, "bank")
    m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
    assert gpu_ctx.check_linearizable(h, m)["valid"] == H.VALID
    assert gpu_ctx.check_bank_totals(h, m, 0)["valid"] == oracle_mod.check_bank_totals(h, m, 0)["valid"] == H.VALID
    # a generated ledger history written as EDN, read back, checked on the device: a torn read must be found
    led = []
    idx = 0
    def emit(type_, process, value):
        nonlocal idx
        led.append({"type": type_, "f": "txn", "value": value, "process": process, "time": 10 * idx, "index": idx})
        idx += 1
    emit("invoke", 0, [["t", 1, {"debit-acct": 1, "credit-acct": 2, "amount": 4}]])
    emit("ok", 0, [["t", 1, {"debit-acct": 1, "credit-acct": 2, "amount": 4}]])
    emit("invoke", 1, [["r", 1, None], ["r", 2, None]])
    emit("ok", 1, [["r", 1, {"credits-posted": 0, "debits-posted": 4}], ["r", 2, {"credits-posted": 0, "debits-posted": 0}]])   # torn
    text = _edn_of(led)
    h = H.flatten_ops(edn.read_history(text), "bank")
    g, o = gpu_ctx.check_linearizable(h, m), oracle_mod.check_linearizable(h, m, 3, eager_reads=True)
    assert g["valid"] == o["valid"] == H.INVALID
    assert g["shards"][0]["witness_index"] == o["shards"][0]["witness_index"] == 3
    assert gpu_ctx.check_bank_totals(h, m, 0)["valid"] == oracle_mod.check_bank_totals(h, m, 0)["valid"] == H.INVALID
