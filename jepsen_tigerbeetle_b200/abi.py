"""ctypes images of the result/option structs in include/jtb_check.h (keep in lock-step)."""
from __future__ import annotations

import ctypes as C

from .history import MAX_ACCOUNTS  # noqa: F401  (re-exported for convenience)

ABI_VERSION = 1

CAUSE_NONE, CAUSE_TABLE_FULL, CAUSE_BUDGET, CAUSE_TOO_WIDE = 0, 1, 2, 3
CAUSE_NAME = {0: None, 1: "table-full", 2: "budget", 3: "too-wide"}
SF_NEVER_READ, SF_STABLE, SF_LOST = 0, 1, 2
BANK_OK, BANK_UNEXPECTED_KEY, BANK_NIL_BALANCE, BANK_WRONG_TOTAL, BANK_NEGATIVE_VALUE = range(5)
BANK_ERR_NAME = {1: "unexpected-key", 2: "nil-balance", 3: "wrong-total", 4: "negative-value"}


class COpts(C.Structure):
    _fields_ = [("device", C.c_int32), ("reserved0", C.c_int32), ("table_bytes", C.c_uint64),
                ("max_configs", C.c_uint64), ("time_budget_ms", C.c_uint32),
                ("search_ctas", C.c_uint32)]


class CLinShard(C.Structure):
    _fields_ = [("valid", C.c_int32), ("witness_index", C.c_int32),
                ("previous_ok_index", C.c_int32), ("cause", C.c_int32),
                ("configs_explored", C.c_uint64), ("probes", C.c_uint64)]


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
                ("reserved0", C.c_int32), ("stable_latency_max_ms", C.c_int64),
                ("lost_latency_max_ms", C.c_int64)]


class CSetFullOut(C.Structure):
    _fields_ = [("shards", C.c_void_p), ("elem_capacity", C.c_int64), ("elem_off", C.c_void_p),
                ("elem_id", C.c_void_p), ("elem_outcome", C.c_void_p),
                ("elem_latency_ms", C.c_void_p), ("elem_dup_count", C.c_void_p),
                ("valid", C.c_int32), ("n_failures", C.c_int32), ("seconds_kernel", C.c_double),
                ("seconds_total", C.c_double)]


class CBankResult(C.Structure):
    _fields_ = [("valid", C.c_int32), ("reserved0", C.c_int32), ("read_count", C.c_int64),
                ("error_count", C.c_int64), ("first_error_index", C.c_int32),
                ("first_error_type", C.c_int32), ("count_by_type", C.c_int64 * 5),
                ("first_index_by_type", C.c_int32 * 5), ("last_index_by_type", C.c_int32 * 5),
                ("worst_index_by_type", C.c_int32 * 5), ("lowest_total", C.c_int64),
                ("highest_total", C.c_int64), ("lowest_index", C.c_int32),
                ("highest_index", C.c_int32), ("seconds_kernel", C.c_double),
                ("seconds_total", C.c_double)]
