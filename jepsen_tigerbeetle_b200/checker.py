"""Host-side mirror of the reference's checker interface (the drop-in boundary).

The reference composes `jepsen.checker/Checker`s — one method, `(check [this test history opts])`
returning a map with a mandatory `:valid?` in {true, false, :unknown}:

    src/tigerbeetle/workloads/set_full.clj:155-158
        (independent/checker (checker/compose {:set-full (checker/set-full {:linearizable? true})
                                               :read-all-invoked-adds (read-all-invoked-adds)}))
    src/tigerbeetle/tests/ledger.clj:363-367
        (checker/compose {:SI (checker opts) :plot ... :lookup-transfers ... :final-reads ... :unexpected-ops ...})

This module gives the same names, argument meaning and error behaviour in Python, backed ONLY by the
CUDA library (native.Context -> libjtb_check.so).  Result maps use the Jepsen key names without the
leading colon ("valid?", "lost-count", ...); `:unknown` is the string "unknown".

Histories may be given as Jepsen op maps (list of dicts) or as an already flattened FlatHistory.
"""
from __future__ import annotations

import traceback
from typing import Any, Mapping

import numpy as np

from . import abi
from .history import (INVALID, MODEL_BANK, MODEL_CAS_REGISTER, MODEL_REGISTER, MODEL_SET, UNKNOWN,
                      VALID, VERDICT_NAME, FlatHistory, flatten_ops, make_model, merge_valid)
from .native import Context

_MODEL_KIND = {"register": MODEL_REGISTER, "cas-register": MODEL_CAS_REGISTER, "set": MODEL_SET,
               "bank": MODEL_BANK}
_CODE = {True: VALID, "unknown": UNKNOWN, False: INVALID}


def valid_code(v: Any) -> int:
    """:valid? value -> lattice code (true < :unknown < false)."""
    return _CODE[v if v in (True, False) else "unknown"]


def merge_valid_values(vals) -> Any:
    """jepsen.checker/merge-valid over :valid? values."""
    return VERDICT_NAME[merge_valid(valid_code(v) for v in vals)]


class Checker:
    """`jepsen.checker/Checker` protocol."""

    def check(self, test: Mapping[str, Any], history, opts: Mapping[str, Any] | None = None) -> dict:
        raise NotImplementedError


def check_safe(checker: Checker, test, history, opts=None) -> dict:
    """jepsen.checker/check-safe: a checker that throws yields {:valid? :unknown :error ...}."""
    try:
        return checker.check(test, history, opts or {})
    except Exception:  # noqa: BLE001 - mirrors (catch Throwable t ...)
        return {"valid?": "unknown", "error": traceback.format_exc()}


def _flat(history, model: str) -> FlatHistory:
    if isinstance(history, FlatHistory):
        return history
    return flatten_ops(history, model)


class _Native:
    """Shares one native context per device between the checkers of a compose map."""

    def __init__(self, ctx: Context | None = None, device: int = 0, **ctx_opts) -> None:
        self._ctx = ctx
        self._device = device
        self._opts = ctx_opts

    @property
    def ctx(self) -> Context:
        if self._ctx is None:
            self._ctx = Context(device=self._device, **self._opts)
        return self._ctx


class Linearizable(Checker, _Native):
    """`(checker/linearizable {:model m})` — knossos analysis on the GPU (hot path A9).

    Result keys follow knossos: valid?, op (witness :index), previous-ok, configs (the configurations stuck at the
    witness, first 10 as jepsen.checker/linearizable keeps them, plus configs-total), configs-explored, analyzer,
    cause (when :unknown).  For keyed histories use `independent_checker(linearizable(...))`."""

    def __init__(self, model: str, ctx: Context | None = None, init_value=None, **ctx_opts) -> None:
        _Native.__init__(self, ctx, **ctx_opts)
        if model not in _MODEL_KIND:
            raise AssertionError("The linearizable checker requires a model")  # as upstream asserts
        self.model = model
        self.init_value = init_value

    def _cmodel(self, test):
        kind = _MODEL_KIND[self.model]
        if kind == MODEL_BANK:
            accounts = list(test.get("accounts", range(1, 9)))
            return make_model(kind, accounts=accounts,
                              init_balance=test.get("initial-balances"),
                              negative_balances_ok=bool(test.get("negative-balances?", True)))
        from .history import NIL
        return make_model(kind, init_value=NIL if self.init_value is None else int(self.init_value))

    def _render_config(self, cm, c: dict) -> dict:
        from .history import NIL
        kind = _MODEL_KIND[self.model]
        if kind == MODEL_BANK:
            model = {int(cm.account_ids[i]): c["balances"][i] for i in range(cm.n_accounts)}
        elif kind == MODEL_SET:
            model = None   # the set is the union of the linearized adds
        else:
            model = None if c["state"] == NIL else c["state"]
        return {"model": model, "pending": [{"index": i} for i in c["pending"]],
                "linearized-open": [{"index": i} for i in c["linearized_open"]],
                "crashed-linearized": c["crashed_linearized"]}

    def check_flat(self, test, h: FlatHistory) -> tuple[dict, list[dict]]:
        cm = self._cmodel(test)
        r = self.ctx.check_linearizable(h, cm)
        per = []
        for k, s in enumerate(r["shards"]):
            m = {"valid?": VERDICT_NAME[s["valid"]], "analyzer": "wgl-gpu"}
            if s["valid"] == INVALID:
                m["op"] = {"index": s["witness_index"]}
                m["previous-ok"] = ({"index": s["previous_ok_index"]}
                                    if s["previous_ok_index"] >= 0 else None)
                fc = self.ctx.final_configs(h, cm, shard=k, cap=10)   # reads the table of the search just done
                m["configs"] = [self._render_config(cm, c) for c in fc["configs"]]
                m["configs-total"] = fc["total"]
            if s["valid"] == UNKNOWN:
                m["cause"] = abi.CAUSE_NAME.get(s["cause"], "unknown")
            per.append(m)
        top = {"valid?": VERDICT_NAME[r["valid"]], "configs-explored": r["configs"],
               "probes": r["probes"], "seconds-kernel": r["seconds_kernel"],
               "seconds-total": r["seconds_total"]}
        return top, per

    def check(self, test, history, opts=None) -> dict:
        h = _flat(history, self.model)
        if h.n_shards != 1:
            raise ValueError("history has independent keys: wrap with independent_checker(...)")
        top, per = self.check_flat(test, h)
        out = dict(per[0])
        out.update({k: v for k, v in top.items() if k != "valid?"})
        return out


class SetFull(Checker, _Native):
    """`(checker/set-full {:linearizable? L})` as called at workloads/set_full.clj:157 (hot path A4)."""

    def __init__(self, checker_opts: Mapping[str, Any] | None = None, ctx: Context | None = None,
                 **ctx_opts) -> None:
        _Native.__init__(self, ctx, **ctx_opts)
        checker_opts = checker_opts or {}
        self.linearizable = bool(checker_opts.get("linearizable?", False))

    @staticmethod
    def shard_maps(r: dict) -> list[dict]:
        out = []
        off = r["elem_off"]
        for s, sh in enumerate(r["shards"]):
            lo, hi = int(off[s]), int(off[s + 1])
            ids, oc = r["elem_id"][lo:hi], r["elem_outcome"][lo:hi]
            lat, dup = r["elem_latency_ms"][lo:hi], r["elem_dup_count"][lo:hi]
            stable_lat = np.sort(lat[oc == abi.SF_STABLE])
            lost_lat = np.sort(lat[oc == abi.SF_LOST])

            def quantiles(x):
                if x.size == 0:
                    return None
                # jepsen.checker/frequency-distribution: (nth sorted (min (dec n) (floor (* n p))))
                return {q: int(x[min(x.size - 1, int(np.floor(x.size * q)))]) for q in (0, 0.5, 0.95, 0.99, 1)}

            stale = [(int(i), int(l)) for i, l, o in zip(ids, lat, oc) if o == abi.SF_STABLE and l > 0]
            stale.sort(key=lambda t: -t[1])
            out.append({
                "valid?": VERDICT_NAME[sh["valid"]],
                "attempt-count": sh["attempt_count"], "stable-count": sh["stable_count"],
                "lost-count": sh["lost_count"], "lost": sorted(int(i) for i in ids[oc == abi.SF_LOST]),
                "never-read-count": sh["never_read_count"],
                "never-read": sorted(int(i) for i in ids[oc == abi.SF_NEVER_READ]),
                "stale-count": sh["stale_count"], "stale": sorted(i for i, _ in stale),
                "worst-stale": [{"element": i, "stable-latency": l} for i, l in stale[:8]],
                "stable-latencies": quantiles(stable_lat), "lost-latencies": quantiles(lost_lat),
                "duplicated-count": sh["duplicated_count"],
                "duplicated": {int(i): int(d) for i, d in zip(ids, dup) if d > 1},
            })
        return out

    def check_flat(self, test, h: FlatHistory) -> tuple[dict, list[dict]]:
        r = self.ctx.check_set_full(h, self.linearizable)
        return {"valid?": VERDICT_NAME[r["valid"]], "seconds-kernel": r["seconds_kernel"]}, self.shard_maps(r)

    def check(self, test, history, opts=None) -> dict:
        """Un-keyed history (what `independent/checker` hands its inner checker per key, set_full.clj:155-157)."""
        h = _flat(history, "set")
        if h.n_shards != 1:
            raise ValueError("history has independent keys: wrap with independent_checker(...)")
        return self.check_flat(test, h)[1][0]


class ReadAllInvokedAdds(Checker, _Native):
    """`(read-all-invoked-adds)` — workloads/set_full.clj:51-75: every :final? :ok read must contain every
    :add value invoked in its sub-history; else {:valid? false :suspect-final-reads [[index missing] ...]}.
    Evaluated on the device in the same pass as set-full (shares its read x element bit-matrix)."""

    def __init__(self, ctx: Context | None = None, **ctx_opts) -> None:
        _Native.__init__(self, ctx, **ctx_opts)

    def check_flat(self, test, h: FlatHistory) -> tuple[dict, list[dict]]:
        r = self.ctx.check_set_full(h, True)
        per = [{"valid?": True} for _ in range(h.n_shards)]
        for sus in r["suspect_final_reads"]:
            m = per[sus["shard"]]
            m["valid?"] = False
            m.setdefault("suspect-final-reads", []).append([sus["index"], sorted(sus["missing"])])
        return {"valid?": r["raia_valid"] == VALID}, per

    def check(self, test, history, opts=None) -> dict:
        h = _flat(history, "set")
        if h.n_shards != 1:
            raise ValueError("history has independent keys: wrap with independent_checker(...)")
        return self.check_flat(test, h)[1][0]


class BankTotals(Checker, _Native):
    """The ledger test's `:SI` checker (tests/ledger.clj:154-192): every :ok read must sum to
    (:total-amount test); unless :negative-balances? no balance may be negative."""

    def __init__(self, checker_opts: Mapping[str, Any] | None = None, ctx: Context | None = None,
                 **ctx_opts) -> None:
        _Native.__init__(self, ctx, **ctx_opts)
        self.negative_balances = bool((checker_opts or {}).get("negative-balances?", False))

    def check(self, test, history, opts=None) -> dict:
        h = _flat(history, "bank")
        m = make_model(MODEL_BANK, accounts=list(test.get("accounts", range(1, 9))),
                       negative_balances_ok=self.negative_balances)
        r = self.ctx.check_bank_totals(h, m, int(test.get("total-amount", 0)))
        errors = {}
        for t, name in abi.BANK_ERR_NAME.items():
            if r["count_by_type"][t]:
                e = {"count": r["count_by_type"][t], "first": {"op": {"index": r["first_index_by_type"][t]}},
                     "worst": {"op": {"index": r["worst_index_by_type"][t]}},
                     "last": {"op": {"index": r["last_index_by_type"][t]}}}
                if name == "wrong-total":
                    e["lowest"] = {"total": r["lowest_total"], "op": {"index": r["lowest_index"]}}
                    e["highest"] = {"total": r["highest_total"], "op": {"index": r["highest_index"]}}
                errors[name] = e
        first = None
        if r["error_count"]:
            first = {"type": abi.BANK_ERR_NAME[r["first_error_type"]], "op": {"index": r["first_error_index"]}}
        out = {"valid?": r["error_count"] == 0, "read-count": r["read_count"],
               "error-count": r["error_count"], "first-error": first, "errors": errors}
        if r["reference_throws"]:
            # tests/ledger.clj:122-123: err-badness divides by (:total-amount test) = 0 (the default, :356) as soon as
            # util/max-by compares two :wrong-total errors; the reference checker throws and jepsen's check-safe
            # reports :unknown.  Same verdict here; the statistics are kept as extra keys.
            out["valid?"] = "unknown"
            out["error"] = "java.lang.ArithmeticException: Divide by zero (err-badness, tests/ledger.clj:122-123)"
        return out


class Compose(Checker):
    """`(checker/compose {name checker ...})`: run each, `:valid?` = merge-valid of the results."""

    def __init__(self, checkers: Mapping[str, Checker]) -> None:
        self.checkers = dict(checkers)

    def check(self, test, history, opts=None) -> dict:
        out = {k: check_safe(c, test, history, opts) for k, c in self.checkers.items()}
        out["valid?"] = merge_valid_values(r["valid?"] for r in list(out.values()))
        return out


class Independent(Checker):
    """`(independent/checker inner)`: split the history by key, check every sub-history, merge
    (SURVEY A.2).  Unlike the JVM original the split is the CSR partition of the flattened history
    and keyed-aware inner checkers (Linearizable, SetFull, Compose of those) receive ALL shards in one
    native call, so the GPU owns the fan-out."""

    def __init__(self, inner: Checker, model: str | None = None) -> None:
        self.inner = inner
        self.model = model

    def _model(self) -> str:
        if self.model:
            return self.model
        c = self.inner
        if isinstance(c, Compose):
            c = next(iter(c.checkers.values()))
        if isinstance(c, Linearizable):
            return c.model
        return "set"

    def _per_key(self, checker: Checker, test, h: FlatHistory, opts) -> list[dict]:
        if isinstance(checker, (Linearizable, SetFull, ReadAllInvokedAdds)):
            try:
                return checker.check_flat(test, h)[1]
            except Exception:  # noqa: BLE001
                err = traceback.format_exc()
                return [{"valid?": "unknown", "error": err} for _ in range(h.n_shards)]
        if isinstance(checker, Compose):
            cols = {name: self._per_key(c, test, h, opts) for name, c in checker.checkers.items()}
            out = []
            for s in range(h.n_shards):
                m = {name: col[s] for name, col in cols.items()}
                m["valid?"] = merge_valid_values(r["valid?"] for r in list(m.values()))
                out.append(m)
            return out
        # generic checker: one call per sub-history
        return [check_safe(checker, test, h.shard(s),
                           dict(opts or {}, **{"history-key": int(h.key_ids[s])}))
                for s in range(h.n_shards)]

    def check(self, test, history, opts=None) -> dict:
        h = _flat(history, self._model())
        per = self._per_key(self.inner, test, h, opts)
        results = {int(k): r for k, r in zip(h.key_ids, per)}
        failures = [k for k, r in results.items() if r["valid?"] is not True]
        return {"valid?": merge_valid_values(r["valid?"] for r in per) if per else True,
                "results": results, "failures": failures}


# ---- constructors with the reference's names ---------------------------------------------------------
def linearizable(opts: Mapping[str, Any], **kw) -> Linearizable:
    """`(checker/linearizable {:model :cas-register})`"""
    return Linearizable(opts["model"], init_value=opts.get("init-value"), **kw)


def set_full(opts: Mapping[str, Any] | None = None, **kw) -> SetFull:
    """`(checker/set-full {:linearizable? true})` — set_full.clj:157"""
    return SetFull(opts, **kw)


def read_all_invoked_adds(**kw) -> ReadAllInvokedAdds:
    """`(read-all-invoked-adds)` — set_full.clj:51-75, :158"""
    return ReadAllInvokedAdds(**kw)


def bank_checker(opts: Mapping[str, Any] | None = None, **kw) -> BankTotals:
    """`(ledger/checker {:negative-balances? true})` — tests/ledger.clj:154-192, :363"""
    return BankTotals(opts, **kw)


def compose(checkers: Mapping[str, Checker]) -> Compose:
    return Compose(checkers)


def independent_checker(inner: Checker, model: str | None = None) -> Independent:
    return Independent(inner, model)


# =====================================================================================================
# The ledger test's remaining (host-side, O(events)) checkers — SURVEY §8(f) N3.  In the reference these
# are plain Clojure seq operations over the op maps with no arithmetic worth a kernel; they are restated
# here on the host so that the whole `compose` map of tests/ledger.clj:363-367 can be served by this
# package.  They consume op maps (not the flattened arrays: they need the :l-t ops and :final? flags).
# =====================================================================================================
def _is_client(op) -> bool:
    p = op.get("process", op.get(":process"))
    return isinstance(p, (int, np.integer)) and not isinstance(p, bool)


def _g(op, name, default=None):
    return op.get(name, op.get(":" + name, default))


def _kw(x):
    return x[1:] if isinstance(x, str) and x.startswith(":") else x


def _txn_f(op):
    """tests/ledger.clj:17-21 op->txn-f: the first micro-op's tag."""
    v = _g(op, "value")
    if not v:
        return None
    first = v[0] if isinstance(v, (list, tuple)) else None
    return _kw(first[0]) if first else None


def _freeze(x):
    if isinstance(x, dict):
        return tuple(sorted((k, _freeze(v)) for k, v in x.items()))
    if isinstance(x, (list, tuple)):
        return tuple(_freeze(v) for v in x)
    if isinstance(x, (set, frozenset)):
        return frozenset(_freeze(v) for v in x)
    return x


class UnexpectedOps(Checker):
    """`(unexpected-ops)` — tests/ledger.clj:194-220: never-resolved invokes and :fail ops mark the
    result :unknown ({:valid? :unknown :open-ops [...] :fail-ops [...]})."""

    def check(self, test, history, opts=None) -> dict:
        hist = [op for op in history if _is_client(op)]
        end_time = _g(hist[-1], "time", 0) if hist else 0
        open_by_process: dict = {}
        for op in hist:  # knossos.history/unmatched-invokes
            t = _kw(_g(op, "type"))
            p = _g(op, "process")
            if t == "invoke":
                open_by_process[p] = op
            else:
                open_by_process.pop(p, None)
        open_ops = sorted(open_by_process.values(), key=lambda o: _g(o, "index", 0))
        opens = [[(end_time - _g(o, "time", 0)) / 1e6, o] for o in open_ops][::-1]  # util/nanos->ms, rseq
        fails = [op for op in hist if _kw(_g(op, "type")) == "fail"]
        out: dict = {"valid?": True}
        if opens:
            out.update({"valid?": "unknown", "open-ops": opens})
        if fails:
            out.update({"valid?": "unknown", "fail-ops": fails})
        return out


class LookupAllInvokedTransfers(Checker):
    """`(lookup-all-invoked-transfers)` — tests/ledger.clj:222-252: every :final? :ok :l-t lookup must
    contain the id of every invoked transfer."""

    def check(self, test, history, opts=None) -> dict:
        hist = [op for op in history if _is_client(op)]
        invoked = set()
        for op in hist:
            if _txn_f(op) == "t" and _kw(_g(op, "type")) == "invoke":
                for micro in _g(op, "value"):
                    invoked.add(micro[1])
        suspects = []
        for op in hist:
            if _txn_f(op) == "l-t" and _kw(_g(op, "type")) == "ok" and _g(op, "final?"):
                ids = {micro[1] for micro in _g(op, "value")}
                if invoked - ids:
                    suspects.append(op)
        out: dict = {"valid?": True}
        if suspects:
            out.update({"valid?": False, "suspect-final-lookups": suspects})
        return out


class FinalReads(Checker):
    """`(final-reads)` — tests/ledger.clj:254-282: final reads (and final :l-t lookups) must exist and be
    all equal: exactly one distinct :value among the :final? :ok :r ops, and among the :l-t ones."""

    def check(self, test, history, opts=None) -> dict:
        hist = [op for op in history if _is_client(op)]

        def finals(tag):
            return {_freeze(_g(op, "value")) for op in hist
                    if _txn_f(op) == tag and _kw(_g(op, "type")) == "ok" and _g(op, "final?")}

        reads, lookups = finals("r"), finals("l-t")
        out: dict = {"valid?": True}
        if len(reads) != 1:
            out.update({"valid?": False, "unequal-final-reads": reads})
        if len(lookups) != 1:
            out.update({"valid?": False, "unequal-final-lookups": lookups})
        return out


class Stats(Checker):
    """`(checker/stats)` as composed at core.clj:144 [UPSTREAM-RECALL, jepsen.checker/stats]: success / failure counts of
    the client completions, overall and by :f; valid only when every :f has at least one :ok completion
    ({:valid? false} otherwise: an operation that never once succeeded makes the rest of the analysis vacuous).
    Host code, as in the reference (a linear pass over the history; nothing to put on a GPU)."""

    @staticmethod
    def _tally(ops) -> dict:
        ok = sum(1 for o in ops if _kw(_g(o, "type")) == "ok")
        fail = sum(1 for o in ops if _kw(_g(o, "type")) == "fail")
        info = sum(1 for o in ops if _kw(_g(o, "type")) == "info")
        return {"valid?": ok > 0, "count": ok + fail + info, "ok-count": ok, "fail-count": fail, "info-count": info}

    def check(self, test, history, opts=None) -> dict:
        done = [op for op in history if _is_client(op) and _kw(_g(op, "type")) != "invoke"]
        by_f: dict = {}
        for op in done:
            by_f.setdefault(_kw(_g(op, "f")), []).append(op)
        out = self._tally(done)
        out["by-f"] = {f: self._tally(ops) for f, ops in sorted(by_f.items(), key=lambda kv: str(kv[0]))}
        out["valid?"] = merge_valid_bool([r["valid?"] for r in out["by-f"].values()])
        return out


def merge_valid_bool(vs) -> Any:
    """jepsen.checker/merge-valid over {True, "unknown", False} values (True for an empty collection)."""
    order = {True: 0, "unknown": 1, False: 2}
    worst = True
    for v in vs:
        if order[v] > order[worst]:
            worst = v
    return worst


def stats() -> Stats:
    return Stats()


def unexpected_ops() -> UnexpectedOps:
    return UnexpectedOps()


def lookup_all_invoked_transfers() -> LookupAllInvokedTransfers:
    return LookupAllInvokedTransfers()


def final_reads() -> FinalReads:
    return FinalReads()


def ledger_checker(checker_opts: Mapping[str, Any] | None = None, ctx: Context | None = None,
                   linear: bool = True) -> Compose:
    """The ledger test's checker (tests/ledger.clj:363-367) minus the gnuplot plotter, plus the
    linearizability search the north-star adds:
        {:SI (checker opts) :lookup-transfers ... :final-reads ... :unexpected-ops ... [:linear ...]}"""
    cs: dict[str, Checker] = {"SI": bank_checker(checker_opts, ctx=ctx),
                              "lookup-transfers": lookup_all_invoked_transfers(),
                              "final-reads": final_reads(), "unexpected-ops": unexpected_ops()}
    if linear:
        cs["linear"] = linearizable({"model": "bank"}, ctx=ctx)
    return compose(cs)
