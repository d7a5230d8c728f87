"""ctypes images of the result/option structs in include/jtb_check.h (keep in lock-step)."""
from __future__ import annotations

import ctypes as C

from .history import MAX_ACCOUNTS

ABI_VERSION = 2
OPT_NO_EAGER_READS = 1
OPT_NO_SCOUTS = 2
OPT_ENGINE_LEVEL = 4
OPT_ENGINE_WORKLIST = 8
OPT_NO_BEAM = 16

CAUSE_NONE, CAUSE_TABLE_FULL, CAUSE_BUDGET, CAUSE_TOO_WIDE = 0, 1, 2, 3
CAUSE_NAME = {0: None, 1: "table-full", 2: "budget", 3: "too-wide"}
SF_NEVER_READ, SF_STABLE, SF_LOST = 0, 1, 2
BANK_OK, BANK_UNEXPECTED_KEY, BANK_NIL_BALANCE, BANK_WRONG_TOTAL, BANK_NEGATIVE_VALUE = range(5)
BANK_ERR_NAME = {1: "unexpected-key", 2: "nil-balance", 3: "wrong-total", 4: "negative-value"}


class COpts(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_int32), ("table_bytes", C.c_uint64),
                ("max_configs", C.c_uint64), ("time_budget_ms", C.c_uint32),
                ("search_ctas", C.c_uint32)]


class CLinShard(C.Structure):
    _fields_ = [("valid", C.c_int32), ("witness_index", C.c_int32),
                ("previous_ok_index", C.c_int32), ("cause", C.c_int32),
                ("configs_explored", C.c_uint64), ("probes", C.c_uint64)]


class CFinalConfig(C.Structure):
    """jtb_final_config: one of knossos' :configs of an INVALID shard."""
    _fields_ = [("state", C.c_int32), ("balances", C.c_int32 * MAX_ACCOUNTS), ("n_pending", C.c_int32),
                ("n_linearized_open", C.c_int32), ("n_crashed_linearized", C.c_int32),
                ("pending_index", C.c_int32 * 64), ("linearized_open_index", C.c_int32 * 64)]


def final_configs_to_list(buf, n: int) -> list[dict]:
    return [{"state": c.state, "balances": list(c.balances),
             "pending": list(c.pending_index[:c.n_pending]),
             "linearized_open": list(c.linearized_open_index[:c.n_linearized_open]),
             "crashed_linearized": c.n_crashed_linearized} for c in buf[:n]]


class CLinResult(C.Structure):
    _fields_ = [("valid", C.c_int32), ("n_failures", C.c_int32),
                ("configs_explored", C.c_uint64), ("probes", C.c_uint64),
                ("hbm_bytes_algorithmic", C.c_uint64), ("key_bytes", C.c_uint32),
                ("reserved0", C.c_uint32), ("seconds_kernel", C.c_double),
                ("seconds_total", C.c_double)]


class CSetFullShard(C.Structure):
    _fields_ = [("valid", C.c_int32), ("attempt_count", C.c_int32), ("stable_count", C.c_int32),
                ("lost_count", C.c_int32), ("never_read_count", C.c_int32),
                ("stale_count", C.c_int32), ("duplicated_count", C.c_int32),
                ("suspect_final_reads", C.c_int32), ("stable_latency_max_ms", C.c_int64),
                ("lost_latency_max_ms", C.c_int64)]


class CSetFullOut(C.Structure):
    _fields_ = [("shards", C.c_void_p), ("elem_capacity", C.c_int64), ("elem_off", C.c_void_p),
                ("elem_id", C.c_void_p), ("elem_outcome", C.c_void_p),
                ("elem_latency_ms", C.c_void_p), ("elem_dup_count", C.c_void_p),
                ("valid", C.c_int32), ("n_failures", C.c_int32), ("seconds_kernel", C.c_double),
                ("seconds_total", C.c_double),
                ("suspect_capacity", C.c_int64), ("suspect_shard", C.c_void_p),
                ("suspect_index", C.c_void_p), ("suspect_missing_off", C.c_void_p),
                ("missing_capacity", C.c_int64), ("missing_ids", C.c_void_p),
                ("n_suspect", C.c_int64), ("raia_valid", C.c_int32), ("reserved1", C.c_int32)]


class CBankResult(C.Structure):
    _fields_ = [("valid", C.c_int32), ("reference_throws", C.c_int32), ("read_count", C.c_int64),
                ("error_count", C.c_int64), ("first_error_index", C.c_int32),
                ("first_error_type", C.c_int32), ("count_by_type", C.c_int64 * 5),
                ("first_index_by_type", C.c_int32 * 5), ("last_index_by_type", C.c_int32 * 5),
                ("worst_index_by_type", C.c_int32 * 5), ("lowest_total", C.c_int64),
                ("highest_total", C.c_int64), ("lowest_index", C.c_int32),
                ("highest_index", C.c_int32), ("seconds_kernel", C.c_double),
                ("seconds_total", C.c_double)]


def alloc_setfull_out(h, shards):
    """Caller-side buffers for jtb_check_set_full (returns the struct and the numpy arrays backing it)."""
    import numpy as np
    n_add_inv = int(np.count_nonzero((h.f == 3) & (h.type == 0))) + 1
    n_final = int(np.count_nonzero((h.flags & 1) != 0)) + 1
    bufs = {
        "elem_off": np.zeros(h.n_shards + 1, np.int64), "elem_id": np.zeros(n_add_inv, np.int32),
        "elem_outcome": np.zeros(n_add_inv, np.uint8), "elem_latency_ms": np.zeros(n_add_inv, np.int64),
        "elem_dup_count": np.zeros(n_add_inv, np.int32),
        "suspect_shard": np.zeros(n_final, np.int32), "suspect_index": np.zeros(n_final, np.int32),
        "suspect_missing_off": np.zeros(n_final + 1, np.int64),
        "missing_ids": np.zeros(max(1, min(n_final * n_add_inv, 1 << 26)), np.int32),
    }
    out = CSetFullOut()
    out.shards = C.cast(shards, C.c_void_p)
    out.elem_capacity = n_add_inv
    for k in ("elem_off", "elem_id", "elem_outcome", "elem_latency_ms", "elem_dup_count", "suspect_shard",
              "suspect_index", "suspect_missing_off", "missing_ids"):
        setattr(out, k, bufs[k].ctypes.data)
    out.suspect_capacity = n_final
    out.missing_capacity = bufs["missing_ids"].shape[0]
    return out, bufs


SETFULL_SHARD_FIELDS = ("valid", "attempt_count", "stable_count", "lost_count", "never_read_count", "stale_count",
                        "duplicated_count", "suspect_final_reads", "stable_latency_max_ms", "lost_latency_max_ms")


def setfull_to_dict(out, shards, bufs) -> dict:
    n = int(bufs["elem_off"][-1])
    ns = int(out.n_suspect)
    moff = bufs["suspect_missing_off"]
    return {
        "valid": out.valid, "n_failures": out.n_failures, "seconds": out.seconds_total,
        "seconds_kernel": out.seconds_kernel,
        "shards": [{f: getattr(s, f) for f in SETFULL_SHARD_FIELDS} for s in shards],
        "elem_off": bufs["elem_off"].copy(), "elem_id": bufs["elem_id"][:n].copy(),
        "elem_outcome": bufs["elem_outcome"][:n].copy(), "elem_latency_ms": bufs["elem_latency_ms"][:n].copy(),
        "elem_dup_count": bufs["elem_dup_count"][:n].copy(),
        "raia_valid": out.raia_valid,
        "suspect_final_reads": [
            {"shard": int(bufs["suspect_shard"][i]), "index": int(bufs["suspect_index"][i]),
             "missing": [int(x) for x in bufs["missing_ids"][int(moff[i]):int(moff[i + 1])]]} for i in range(ns)],
    }
