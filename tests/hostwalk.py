"""ctypes binding of tests/native/libjtb_hostwalk.so — TEST INFRASTRUCTURE: the product's per-thread expansion core
(csrc/jtb_expand.h, the code the search kernel inlines) and host preparation (csrc/jtb_prep.cpp) compiled for the CPU
and driven by a std::unordered_set, so the device logic is compared with the oracle without a GPU."""
import ctypes as C
import os
import subprocess

from jepsen_tigerbeetle_b200.history import as_c_history

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "native", "libjtb_hostwalk.so")
_SRCS = [os.path.join(_HERE, "native", "hostwalk.cpp"),
         os.path.join(_ROOT, "jepsen_tigerbeetle_b200", "csrc", "jtb_prep.cpp")]
_DEPS = _SRCS + [os.path.join(_ROOT, "jepsen_tigerbeetle_b200", "csrc", f) for f in ("jtb_prep.h", "jtb_expand.h")] + \
        [os.path.join(_ROOT, "include", "jtb_check.h")]
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in _DEPS):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _SO] + _SRCS)
        _lib = C.CDLL(_SO)
    return _lib


def walk(h, model, eager_reads=True, max_configs=0):
    ch = as_c_history(h)
    n = h.n_shards
    valid, wit, prev = (C.c_int32 * n)(), (C.c_int32 * n)(), (C.c_int32 * n)()
    configs, kw = C.c_ulonglong(0), C.c_int32(0)
    rc = lib().jtb_hostwalk(C.byref(ch), C.byref(model), int(eager_reads), C.c_ulonglong(max_configs), valid, wit, prev,
                            C.byref(configs), C.byref(kw))
    if rc != 0:
        raise RuntimeError(f"jtb_hostwalk rc={rc}")
    return {"valid": max(valid) if n else 0, "configs": configs.value, "key_bytes": kw.value * 8,
            "shards": [{"valid": valid[s], "witness_index": wit[s], "previous_ok_index": prev[s]} for s in range(n)]}


def walk_bfs(h, model, eager_reads=True, max_configs=0, width_cap=0):
    """Breadth-first by depth with a visited set that lives for ONE level (the device level engine's rule)."""
    ch = as_c_history(h)
    n = h.n_shards
    valid, wit, prev = (C.c_int32 * n)(), (C.c_int32 * n)(), (C.c_int32 * n)()
    configs, kw, nl = C.c_ulonglong(0), C.c_int32(0), C.c_int(0)
    widths = (C.c_ulonglong * max(1, width_cap))()
    rc = lib().jtb_hostwalk_bfs(C.byref(ch), C.byref(model), int(eager_reads), C.c_ulonglong(max_configs), valid, wit,
                                prev, C.byref(configs), C.byref(kw), widths, width_cap, C.byref(nl))
    if rc != 0:
        raise RuntimeError(f"jtb_hostwalk_bfs rc={rc}")
    return {"valid": max(valid) if n else 0, "configs": configs.value, "key_bytes": kw.value * 8, "levels": nl.value,
            "widths": list(widths[:min(width_cap, nl.value)]),
            "shards": [{"valid": valid[s], "witness_index": wit[s], "previous_ok_index": prev[s]} for s in range(n)]}
