"""ctypes driver of tests/native/libjtb_fakejvm.so — TEST INFRASTRUCTURE: the JNI shim jni/jtb_jni.c compiled against
jni/stub/jni.h and called through a fake JNIEnv (tests/native/fake_jvm.c), i.e. the exact marshalling a JVM would
exercise, in an image without a JDK."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "native", "libjtb_fakejvm.so")
_SRCS = [os.path.join(_HERE, "native", "fake_jvm.c"), os.path.join(_ROOT, "jni", "jtb_jni.c")]
_DEPS = _SRCS + [os.path.join(_ROOT, "jni", "stub", "jni.h"), os.path.join(_ROOT, "include", "jtb_check.h")]
_lib = None


def build():
    from jepsen_tigerbeetle_b200 import native
    libdir = os.path.dirname(native.LIB_PATH)
    if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in _DEPS):
        subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared",
                               "-I" + os.path.join(_ROOT, "jni", "stub"), "-I" + os.path.join(_ROOT, "include"),
                               "-o", _SO] + _SRCS + ["-L" + libdir, "-ljtb_check",
                                                     "-Wl,-rpath,$ORIGIN/../../jepsen_tigerbeetle_b200"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.fj_new_array.restype = C.c_void_p
        L.fj_new_array.argtypes = [C.c_char, C.c_int, C.c_void_p]
        L.fj_free_array.argtypes = [C.c_void_p]
        L.fj_array_len.argtypes = [C.c_void_p]
        L.fj_array_data.restype = C.c_void_p
        L.fj_array_data.argtypes = [C.c_void_p]
        L.fj_exception.restype = C.c_char_p
        L.fj_create.restype = C.c_longlong
        L.fj_create.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_int]
        L.fj_destroy.argtypes = [C.c_longlong]
        L.fj_multi_create.restype = C.c_longlong
        L.fj_multi_create.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_longlong, C.c_int]
        L.fj_multi_destroy.argtypes = [C.c_longlong]
        L.fj_check_linearizable.restype = C.c_void_p
        L.fj_check_linearizable.argtypes = [C.c_longlong, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.fj_final_configs.restype = C.c_void_p
        L.fj_final_configs.argtypes = [C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.fj_check_set_full.restype = C.c_void_p
        L.fj_check_set_full.argtypes = [C.c_longlong, C.c_int, C.c_void_p, C.c_int]
        L.fj_check_bank_totals.restype = C.c_void_p
        L.fj_check_bank_totals.argtypes = [C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int]
        _lib = L
    return _lib


class JavaException(RuntimeError):
    pass


def _jarr(kind, arr):
    arr = np.ascontiguousarray(arr)
    return lib().fj_new_array(kind.encode(), int(arr.size), arr.ctypes.data if arr.size else None)


def jhistory(h):
    """FlatHistory -> the Object[14] jtb.Native takes (fake jarray handles)."""
    elems = [_jarr("b", h.type.astype(np.uint8)), _jarr("b", h.f.astype(np.uint8)), _jarr("b", h.flags.astype(np.uint8)),
             _jarr("i", h.process.astype(np.int32)), _jarr("i", h.index.astype(np.int32)),
             _jarr("l", h.time_ns.astype(np.int64)), _jarr("i", h.a.astype(np.int32)), _jarr("i", h.b.astype(np.int32)),
             _jarr("i", h.c.astype(np.int32)), _jarr("l", h.payload_off.astype(np.int64)),
             _jarr("i", h.payload_len.astype(np.int32)), _jarr("i", h.payload.astype(np.int32)),
             _jarr("l", h.shard_off.astype(np.int64)), _jarr("l", np.asarray(h.key_ids, np.int64))]
    ptrs = (C.c_void_p * 14)(*elems)
    return lib().fj_new_array(b"o", 14, C.addressof(ptrs))


def _result(ptr, dtype):
    L = lib()
    exc = L.fj_exception()
    if exc is not None:
        msg = exc.decode()
        L.fj_clear_exception()
        raise JavaException(msg)
    assert L.fj_outstanding() == 0, "the shim leaked a Get<Type>ArrayElements"
    n = L.fj_array_len(ptr)
    out = np.ctypeslib.as_array(C.cast(L.fj_array_data(ptr), C.POINTER(C.c_int64 if dtype == np.int64 else C.c_int32)),
                                shape=(n,)).copy()
    L.fj_free_array(ptr)
    return out


def create(device=0, flags=0):
    h = lib().fj_create(device, flags, 0, 0, 0)
    _ = lib().fj_exception()
    if _ is not None:
        lib().fj_clear_exception()
        raise JavaException(_.decode())
    return h


def check_linearizable(handle, h, model, multi=False):
    acc = _jarr("i", np.array(list(model.account_ids)[:model.n_accounts], np.int32)) if model.n_accounts else None
    bal = _jarr("i", np.array(list(model.init_balance)[:model.n_accounts], np.int32)) if model.n_accounts else None
    return _result(lib().fj_check_linearizable(handle, int(multi), jhistory(h), model.kind, model.init_value, acc, bal,
                                               model.negative_balances_ok), np.int64)


def final_configs(handle, h, model, shard=0, cap=10):
    acc = _jarr("i", np.array(list(model.account_ids)[:model.n_accounts], np.int32)) if model.n_accounts else None
    bal = _jarr("i", np.array(list(model.init_balance)[:model.n_accounts], np.int32)) if model.n_accounts else None
    return _result(lib().fj_final_configs(handle, jhistory(h), model.kind, model.init_value, acc, bal,
                                          model.negative_balances_ok, shard, cap), np.int32)


def check_set_full(handle, h, linearizable=True, multi=False):
    return _result(lib().fj_check_set_full(handle, int(multi), jhistory(h), int(linearizable)), np.int64)


def check_bank_totals(handle, h, model, total=0):
    acc = _jarr("i", np.array(list(model.account_ids)[:model.n_accounts], np.int32))
    return _result(lib().fj_check_bank_totals(handle, jhistory(h), acc, total, model.negative_balances_ok), np.int64)
