"""The JVM binding (java/jtb/Native.java, jni/jtb_jni.c, clj/jtb/checker.clj) checked without a JVM:
CPU tier — the shim compiles (-Wall -Wextra -Werror) against jni/stub/jni.h and links against libjtb_check.so; every
`native` method of jtb.Native has its Java_jtb_Native_<name> export with the same number of arguments; every
Native/<method> call in the Clojure glue names a declared method with that arity.
GPU tier — the shim is driven end to end through a fake JNIEnv and must return what the direct ctypes binding returns."""
import os
import re
import subprocess

import numpy as np
import pytest

import fakejvm
import kat
from jepsen_tigerbeetle_b200 import history as H
from jepsen_tigerbeetle_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def java_natives():
    src = open(os.path.join(ROOT, "java", "jtb", "Native.java")).read()
    out = {}
    for m in re.finditer(r"public static native\s+[\w\[\]]+\s+(\w+)\s*\(([^)]*)\)", src, re.S):
        args = [a for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = len(args)
    return out


def c_exports():
    src = open(os.path.join(ROOT, "jni", "jtb_jni.c")).read()
    out = {}
    for m in re.finditer(r"JNICALL\s+Java_jtb_Native_(\w+)\s*\(([^)]*)\)", src, re.S):
        out[m.group(1)] = len([a for a in m.group(2).split(",") if a.strip()]) - 2   # minus JNIEnv*, jclass
    return out


def clj_calls():
    """(Native/method arg ...) forms of the glue -> [(method, n_args)] (balanced-paren scan; no strings with parens)."""
    src = open(os.path.join(ROOT, "clj", "jtb", "checker.clj")).read()
    src = re.sub(r";[^\n]*", "", src)
    calls = []
    for m in re.finditer(r"\(Native/(\w+)", src):
        i, depth, n_args, in_tok = m.end(), 0, 0, False
        while True:
            ch = src[i]
            if ch in "([{":
                if depth == 0:
                    n_args += 1
                depth += 1
                in_tok = False
            elif ch in ")]}":
                if depth == 0:
                    break
                depth -= 1
            elif ch.isspace():
                in_tok = False
            elif depth == 0 and not in_tok:
                n_args += 1
                in_tok = True
            i += 1
        calls.append((m.group(1), n_args))
    return calls


def test_shim_compiles_and_links():
    so = fakejvm.build()
    syms = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    for name in java_natives():
        assert f"Java_jtb_Native_{name}" in syms


def test_native_java_matches_jni_exports():
    j, c = java_natives(), c_exports()
    assert j and set(j) == set(c), (sorted(j), sorted(c))
    for name in j:
        assert j[name] == c[name], name


def test_clojure_calls_match_native_java():
    j = java_natives()
    calls = clj_calls()
    assert {m for m, _ in calls} >= {"create", "multiCreate", "checkLinearizable", "finalConfigs", "checkSetFull",
                                    "checkBankTotals"}
    for method, n in calls:
        assert method in j, method
        assert j[method] == n, (method, n, j[method])


def test_glue_covers_every_result_field():
    """The record layouts documented in Native.java are what the shim writes (sizes of the fixed parts)."""
    src = open(os.path.join(ROOT, "jni", "jtb_jni.c")).read()
    assert "8 + 7 * ns" in src and "jlong v[34]" in src and "8 + 10ll * ns" in src


# ---------------------------------------------------------------------------------------------------------------------
def model_for(name):
    if name == "bank":
        return H.make_model(H.MODEL_BANK, accounts=range(1, 9))
    return H.make_model({"register": H.MODEL_REGISTER, "cas-register": H.MODEL_CAS_REGISTER, "set": H.MODEL_SET}[name])


@pytest.mark.gpu
def test_shim_end_to_end_matches_ctypes(gpu_ctx):
    jh = fakejvm.create(0)
    try:
        # linearizable: KATs of every model + a keyed history
        cases = [(H.flatten_ops(kat.ops(text), model), model_for(model)) for _, model, text, _, _ in kat.ALL_LIN_KATS]
        cases.append((synth.generate(synth.SynthSpec("cas-register", 2000, 32, 5, p_info=0.05, n_keys=8,
                                                     grouped_keys=True, stale_read=True)), model_for("cas-register")))
        for h, m in cases:
            g = gpu_ctx.check_linearizable(h, m)
            r = fakejvm.check_linearizable(jh, h, m)
            assert (r[0], r[1], r[7]) == (g["valid"], g["n_failures"], h.n_shards)
            for s, gs in enumerate(g["shards"]):
                rec = r[8 + 7 * s: 15 + 7 * s]
                assert list(rec[:4]) == [gs["valid"], gs["witness_index"], gs["previous_ok_index"], gs["cause"]]
        # :configs of an invalid history
        h = synth.generate(synth.SynthSpec("bank", 400, 6, 3, tau_think_ns=10e6, stale_read=True))
        m = model_for("bank")
        g = gpu_ctx.check_linearizable(h, m)
        assert g["valid"] == H.INVALID
        fc = gpu_ctx.final_configs(h, m, 0, 10)
        assert fakejvm.check_linearizable(jh, h, m)[0] == H.INVALID
        xs = fakejvm.final_configs(jh, h, m, 0, 10)
        assert xs[0] == fc["total"] and len(xs) == 1 + 140 * min(10, fc["total"])
        assert list(xs[2:10]) == fc["configs"][0]["balances"]
        # set-full with detail
        h = synth.config_c4(seed=2, n_keys=4, n_ops=3000)
        g = gpu_ctx.check_set_full(h, True)
        r = fakejvm.check_set_full(jh, h, True)
        ns, ne = int(r[6]), int(r[7])
        assert (r[0], r[1], r[2], ns, ne) == (g["valid"], g["n_failures"], g["raia_valid"], h.n_shards, len(g["elem_id"]))
        for s, gs in enumerate(g["shards"]):
            assert list(r[8 + 10 * s: 18 + 10 * s]) == [gs[f] for f in
                                                         ("valid", "attempt_count", "stable_count", "lost_count",
                                                          "never_read_count", "stale_count", "duplicated_count",
                                                          "suspect_final_reads", "stable_latency_max_ms", "lost_latency_max_ms")]
        e0 = 8 + 10 * ns + ns + 1
        el = r[e0: e0 + 4 * ne].reshape(ne, 4)
        assert np.array_equal(el[:, 0], g["elem_id"]) and np.array_equal(el[:, 1], g["elem_outcome"])
        assert np.array_equal(el[:, 2], g["elem_latency_ms"]) and np.array_equal(el[:, 3], g["elem_dup_count"])
        # bank totals
        h = synth.generate(synth.SynthSpec("bank", 1000, 8, 4, tau_think_ns=20e6))
        h.payload[h.payload_off[np.flatnonzero(h.payload_len > 0)[3]] + 1] += 5
        g = gpu_ctx.check_bank_totals(h, m, 0)
        r = fakejvm.check_bank_totals(jh, h, m, 0)
        assert list(r[:6]) == [g["valid"], g["reference_throws"], g["read_count"], g["error_count"],
                               g["first_error_index"], g["first_error_type"]]
        assert list(r[6:11]) == g["count_by_type"] and list(r[21:26]) == g["worst_index_by_type"]
        # a native error surfaces as a Java exception (check-safe turns it into :unknown)
        bad = H.flatten_ops(kat.ops("0:ok write 1"), "register")
        with pytest.raises(fakejvm.JavaException):
            fakejvm.check_linearizable(jh, bad, model_for("register"))
    finally:
        fakejvm.lib().fj_destroy(jh)
