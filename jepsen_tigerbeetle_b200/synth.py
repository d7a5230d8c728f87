"""Deterministic synthetic Jepsen histories for the BASELINE.json configs.

Counter-based RNG (splitmix64 seeding a PCG32 stream) so the same (config, seed) gives the same
history everywhere.  Per client thread: invoke, Exp(tau_op) duration, a linearization point uniform
inside the op's interval; all linearization points are sorted and applied to the true model to get
return values, so every generated history is linearizable by construction ("valid" variants).
`stale_read` then makes one ok read return an old state (ground truth is decided by the oracle).

Op shapes follow the reference generators:
  set-full   adds/reads 1:1 over random keys      set_full.clj:22-45,159
  bank       reads of all accounts / transfers    tests/ledger.clj:27-67 (amount 1..max-transfer, debit != credit)
  :info      client timeouts                      set_full.clj:107-110 ; workloads/ledger.clj:46-48
             after an :info the thread continues as process + concurrency (jepsen convention)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .history import (F_ADD, F_CAS, F_READ, F_TRANSFER, F_WRITE, NIL, T_FAIL, T_INFO, T_INVOKE,
                      T_OK, FlatHistory)

MASK64 = (1 << 64) - 1


class PCG32:
    """PCG-XSH-RR 64/32 seeded through splitmix64."""

    def __init__(self, seed: int, stream: int = 0) -> None:
        x = (seed * 0x9E3779B97F4A7C15 + stream) & MASK64
        self._sm = x
        self.state = 0
        self.inc = ((self._splitmix() << 1) | 1) & MASK64
        self.state = (self._splitmix() + self.inc) & MASK64
        self.u32()

    def _splitmix(self) -> int:
        self._sm = (self._sm + 0x9E3779B97F4A7C15) & MASK64
        z = self._sm
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def u32(self) -> int:
        old = self.state
        self.state = (old * 6364136223846793005 + self.inc) & MASK64
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def below(self, n: int) -> int:
        """Uniform integer in [0, n) (n < 2**32), multiply-shift."""
        return (self.u32() * n) >> 32

    def unit(self) -> float:
        """Uniform double in (0, 1]."""
        return (self.u32() + 1) / 4294967296.0

    def exp_ns(self, mean_ns: float) -> int:
        return max(2, int(-mean_ns * math.log(self.unit())))


@dataclass
class SynthSpec:
    model: str                   # 'register' | 'cas-register' | 'set' | 'bank'
    n_ops: int
    n_clients: int
    seed: int = 1
    p_info: float = 0.0
    n_keys: int = 1
    grouped_keys: bool = False   # True: threads partitioned over keys (independent/concurrent-generator)
    tau_op_ns: float = 10e6
    tau_think_ns: float = 0.0
    n_values: int = 5            # register values 0..n_values-1
    n_accounts: int = 8          # bank accounts 1..8 (core.clj:208-210)
    max_transfer: int = 5        # tests/ledger.clj:355
    first_element: int = 9       # set_full.clj:159: adds start after the pre-created accounts
    stale_read: bool = False     # make one ok read stale ("invalid" variant; oracle decides)
    stale_frac: float = 0.9
    stale_by: int = 0            # linearization steps back (0 => 8 * clients-per-key)
    final_reads: bool = False    # quiesce + one :final? read per key (set_full.clj:161-170)


def generate(spec: SynthSpec) -> FlatHistory:
    rng = PCG32(spec.seed, 1)
    C, K = spec.n_clients, spec.n_keys
    model = spec.model
    # ---- phase 1: op intervals per thread -------------------------------------------------------
    next_inv = [0] * C
    for t in range(C):
        next_inv[t] = 1 + t  # staggered by 1 ns so the first invokes have a defined order
    proc = list(range(C))
    ops = []  # dicts are slow; use tuples: (t_inv, t_ret, t_lin, thread, process, key, f, a, b, c, fate)
    next_elem = spec.first_element
    import heapq
    heap = [(next_inv[t], t) for t in range(C)]
    heapq.heapify(heap)
    per_key_threads = max(1, C // K) if spec.grouped_keys else C
    for _ in range(spec.n_ops):
        t_inv, t = heapq.heappop(heap)
        dur = rng.exp_ns(spec.tau_op_ns)
        t_ret = t_inv + dur
        t_lin = t_inv + 1 + rng.below(max(1, min(dur - 1, 0xFFFFFFFF)))
        if spec.grouped_keys:
            key = min(K - 1, t // per_key_threads)
        else:
            key = rng.below(K) if K > 1 else 0
        a = b = c = 0
        if model in ("register", "cas-register"):
            kind = rng.below(3 if model == "cas-register" else 2)
            if kind == 0:
                f = F_READ
            elif kind == 1:
                f, a = F_WRITE, rng.below(spec.n_values)
            else:
                f, a, b = F_CAS, rng.below(spec.n_values), rng.below(spec.n_values)
        elif model == "set":
            if rng.below(2) == 0:
                f, a = F_ADD, next_elem
                next_elem += 1
            else:
                f = F_READ
        elif model == "bank":
            if rng.below(2) == 0:
                f = F_READ
            else:
                f = F_TRANSFER
                b = 1 + rng.below(spec.n_accounts)
                c = 1 + rng.below(spec.n_accounts - 1)
                if c >= b:
                    c += 1
                a = 1 + rng.below(spec.max_transfer)
        else:
            raise ValueError(model)
        fate = 0  # 0 ok, 1 info+applied, 2 info+not applied
        if spec.p_info > 0 and rng.unit() <= spec.p_info:
            fate = 1 if rng.below(2) == 0 else 2
        ops.append([t_inv, t_ret, t_lin, t, proc[t], key, f, a, b, c, fate])
        if fate:
            proc[t] += C
        think = rng.exp_ns(spec.tau_think_ns) if spec.tau_think_ns > 0 else 1
        heapq.heappush(heap, (t_ret + think, t))
    # ---- phase 2: apply in linearization order --------------------------------------------------
    order = sorted(range(len(ops)), key=lambda i: (ops[i][2], i))
    reg = [NIL] * K
    bal = [[0] * spec.n_accounts for _ in range(K)]
    sets: list[list[int]] = [[] for _ in range(K)]
    results: list = [None] * len(ops)   # read values / cas success
    snapshots: list[list] = [[] for _ in range(K)]  # per key: state snapshots per lin step (for stale reads)
    lin_pos = [0] * len(ops)
    want_snap = spec.stale_read
    for i in order:
        (_ti, _tr, _tl, _t, _p, key, f, a, b, c, fate) = ops[i]
        if want_snap:
            lin_pos[i] = len(snapshots[key])
            if model in ("register", "cas-register"):
                snapshots[key].append(reg[key])
            elif model == "bank":
                snapshots[key].append(tuple(bal[key]))
            else:
                snapshots[key].append(len(sets[key]))
        if fate == 2:
            continue
        if f == F_READ:
            if model in ("register", "cas-register"):
                results[i] = reg[key]
            elif model == "bank":
                results[i] = tuple(bal[key])
            else:
                results[i] = len(sets[key])  # prefix length of sets[key] in lin order
        elif f == F_WRITE:
            reg[key] = a
        elif f == F_CAS:
            if reg[key] == a:
                reg[key] = b
                results[i] = True
            else:
                results[i] = False
        elif f == F_ADD:
            sets[key].append(a)
        elif f == F_TRANSFER:
            bal[key][b - 1] -= a
            bal[key][c - 1] += a
    # ---- stale-read mutation --------------------------------------------------------------------
    mutated = -1
    if spec.stale_read:
        cand = [i for i in sorted(range(len(ops)), key=lambda i: ops[i][1])
                if ops[i][6] == F_READ and ops[i][10] == 0 and ops[i][5] == 0]
        if cand:
            back = spec.stale_by or 8 * per_key_threads
            start = int(spec.stale_frac * len(cand))
            for i in cand[start:] + cand[:start][::-1]:
                key = ops[i][5]
                p = max(0, lin_pos[i] - back)
                old = snapshots[key][p]
                if old != results[i] and not (model == "set" and old >= results[i]):
                    results[i] = old
                    mutated = i
                    break
    # ---- phase 3: events ------------------------------------------------------------------------
    ev = []  # (time, seq, op, is_completion)
    for i, o in enumerate(ops):
        ev.append((o[0], 0, i, 0))
        ev.append((o[1], 1, i, 1))
    t_end = max(o[1] for o in ops) if ops else 0
    final_ops = []
    if spec.final_reads:
        # quiesce 5 s then one final read per key on thread 0.. (set_full.clj:161-170)
        tq = t_end + 5_000_000_000
        for key in range(K):
            i = len(ops) + len(final_ops)
            final_ops.append([tq + 2 * key, tq + 2 * key + 1, 0, key % C, proc[key % C], key,
                              F_READ, 0, 0, 0, 0])
            ev.append((tq + 2 * key, 0, i, 0))
            ev.append((tq + 2 * key + 1, 1, i, 1))
    ev.sort()
    all_ops = ops + final_ops
    n = len(ev)
    typ = np.zeros(n, np.uint8); f_arr = np.zeros(n, np.uint8); flags = np.zeros(n, np.uint8)
    proc_arr = np.zeros(n, np.int32); idx = np.arange(n, dtype=np.int32)
    time_arr = np.zeros(n, np.int64)
    a_arr = np.zeros(n, np.int32); b_arr = np.zeros(n, np.int32); c_arr = np.zeros(n, np.int32)
    plen = np.zeros(n, np.int32)
    keys_arr = np.zeros(n, np.int64)
    payload_chunks: list[np.ndarray] = [None] * n  # type: ignore[list-item]
    sets_np = [np.array(s, dtype=np.int32) for s in sets]
    empty = np.zeros(0, np.int32)
    acct_ids = np.arange(1, spec.n_accounts + 1, dtype=np.int32)
    for e, (tm, _seq, i, comp) in enumerate(ev):
        o = all_ops[i]
        is_final = i >= len(ops)
        f = o[6]
        time_arr[e] = tm; proc_arr[e] = o[4]; f_arr[e] = f; keys_arr[e] = o[5]
        a_arr[e], b_arr[e], c_arr[e] = o[7], o[8], o[9]
        payload_chunks[e] = empty
        if is_final:
            flags[e] = 1
        if not comp:
            typ[e] = T_INVOKE
            if f == F_READ:
                a_arr[e] = NIL
                plen[e] = -1
            continue
        fate = o[10]
        if fate:
            typ[e] = T_INFO
            if f == F_READ:
                a_arr[e] = NIL
                plen[e] = -1
            continue
        typ[e] = T_OK
        if f == F_CAS and not results[i]:
            typ[e] = T_FAIL
        elif f == F_READ:
            if model in ("register", "cas-register"):
                a_arr[e] = results[i]
            elif model == "bank":
                r = results[i]
                pl = np.empty(2 * spec.n_accounts, np.int32)
                pl[0::2] = acct_ids
                pl[1::2] = r
                payload_chunks[e] = pl
                plen[e] = pl.shape[0]
            else:
                k = o[5]
                cnt = len(sets[k]) if is_final else results[i]
                pl = np.sort(sets_np[k][:cnt])
                payload_chunks[e] = pl
                plen[e] = cnt
    # ---- CSR by key ------------------------------------------------------------------------------
    perm = np.argsort(keys_arr, kind="stable")
    counts = np.bincount(keys_arr.astype(np.int64), minlength=K)
    shard_off = np.zeros(K + 1, np.int64)
    np.cumsum(counts, out=shard_off[1:])
    plen_p = plen[perm]
    lens = np.maximum(plen_p, 0).astype(np.int64)
    poff = np.zeros(n, np.int64)
    if n:
        np.cumsum(lens[:-1], out=poff[1:])
    chunks = [payload_chunks[j] for j in perm]
    payload = np.concatenate(chunks).astype(np.int32) if chunks else empty
    meta = {"model": model, "spec": spec, "mutated_op_index": mutated,
            "n_ops": len(all_ops), "accounts": list(range(1, spec.n_accounts + 1))}
    h = FlatHistory(typ[perm], f_arr[perm], flags[perm], proc_arr[perm], idx[perm], time_arr[perm],
                    a_arr[perm], b_arr[perm], c_arr[perm], poff, plen_p, payload, shard_off,
                    np.arange(1, K + 1, dtype=np.int64), meta)
    h.validate()
    return h


# ---- the five BASELINE.json configs ---------------------------------------------
def config_c1(seed: int = 1, **kw) -> FlatHistory:
    """set-full :linearizable? true, 100 ops, 4 clients, 1 key."""
    return generate(SynthSpec("set", 100, 4, seed, final_reads=True, **kw))


def config_c2(seed: int = 1, p_info: float = 0.0, **kw) -> FlatHistory:
    """1k-op cas-register history, 16 concurrent clients."""
    return generate(SynthSpec("cas-register", 1000, 16, seed, p_info=p_info, **kw))


def config_c3(seed: int = 1, p_info: float = 0.0, **kw) -> FlatHistory:
    """10k-op bank-transfer history, 32 clients (headline)."""
    return generate(SynthSpec("bank", 10000, 32, seed, p_info=p_info, **kw))


def config_c4(seed: int = 1, n_keys: int = 64, p_info: float = 0.01, n_ops: int = 100000,
              **kw) -> FlatHistory:
    """100k-op set-full grow-only history, 64 clients, K ledgers."""
    return generate(SynthSpec("set", n_ops, 64, seed, p_info=p_info, n_keys=n_keys,
                              final_reads=True, **kw))


def config_c5(seed: int = 1, n_keys: int = 256, p_info: float = 0.30, n_ops: int = 50000,
              **kw) -> FlatHistory:
    """50k-op adversarial cas-register history, 30% :info, 8 clients per key."""
    return generate(SynthSpec("cas-register", n_ops, 8 * n_keys, seed, p_info=p_info,
                              n_keys=n_keys, grouped_keys=True, **kw))


def poison_c5(h: FlatHistory, shard: int = 7) -> FlatHistory:
    """One key of a C5 history made non-linearizable: its third :ok read returns a value nobody ever wrote
    (values are 0..4).  The merged verdict of the keyed check must flip to invalid with exactly one failure."""
    import copy
    from . import history as H
    h = copy.deepcopy(h)
    lo, hi = int(h.shard_off[shard]), int(h.shard_off[shard + 1])
    reads = [e for e in range(lo, hi) if h.f[e] == H.F_READ and h.type[e] == H.T_OK]
    h.a[reads[min(2, len(reads) - 1)]] = 99
    return h


def poison_c4(h: FlatHistory, shard: int = 5) -> FlatHistory:
    """One ledger of a C4 history loses elements: the two largest ids vanish from the ledger's LAST non-final :ok read
    (payloads are sorted), so elements a read has already seen are :lost -> set-full invalid for that ledger."""
    import copy
    from . import history as H
    h = copy.deepcopy(h)
    lo, hi = int(h.shard_off[shard]), int(h.shard_off[shard + 1])
    reads = [e for e in range(lo, hi) if h.f[e] == H.F_READ and h.type[e] == H.T_OK and h.payload_len[e] > 4
             and not (h.flags[e] & 1)]
    h.payload_len[reads[-1]] -= 2
    return h
