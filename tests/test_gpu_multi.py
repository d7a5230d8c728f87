"""The in-library multi-GPU fan-out (`jtb_multi_*`, csrc/jtb_multi.cpp: LPT partition of the keys over the devices,
one context + host thread per device, verdict arrays merged with one NCCL allReduce(max)) against the single-context
result on the same history.  Runs with however many GPUs the box has (1 GPU: the same code path without NCCL)."""
import numpy as np
import pytest

from jepsen_tigerbeetle_b200 import history as H
from jepsen_tigerbeetle_b200 import synth

pytestmark = pytest.mark.gpu


def test_multi_matches_single_context_on_the_sharded_bench_histories(gpu_ctx):
    from jepsen_tigerbeetle_b200 import native
    c5 = synth.poison_c5(synth.config_c5(seed=1, n_ops=20000, n_keys=96), 7)
    c4 = synth.poison_c4(synth.config_c4(seed=1, n_keys=24, n_ops=30000), 5)
    mc = H.make_model(H.MODEL_CAS_REGISTER)
    with native.MultiContext(n_gpus=0) as mctx:
        assert mctx.n_gpus == native.device_count()
        g = mctx.check_linearizable(c5, mc)
        s = gpu_ctx.check_linearizable(c5, mc)
        assert (g["valid"], g["n_failures"]) == (s["valid"], s["n_failures"]) == (H.INVALID, 1)
        for a, b in zip(g["shards"], s["shards"]):
            assert (a["valid"], a["witness_index"], a["previous_ok_index"]) == (b["valid"], b["witness_index"], b["previous_ok_index"])
        assert set(g["device_of_shard"]) == set(range(mctx.n_gpus))       # every device got keys
        gs = mctx.check_set_full(c4, True)
        ss = gpu_ctx.check_set_full(c4, True)
        assert (gs["valid"], gs["n_failures"]) == (ss["valid"], ss["n_failures"]) == (H.INVALID, 1)
        assert gs["shards"] == ss["shards"]


def test_multi_two_bank_ledgers(gpu_ctx):
    """Two independent bank ledgers (keys), one with a stale read: one key per device when there are two."""
    from jepsen_tigerbeetle_b200 import native
    parts = [synth.generate(synth.SynthSpec("bank", 3000, 32, 2, tau_think_ns=20e6, stale_read=True)),
             synth.generate(synth.SynthSpec("bank", 3000, 32, 3, tau_think_ns=20e6))]
    h = H.concat_keys(parts)
    m = H.make_model(H.MODEL_BANK, accounts=range(1, 9))
    with native.MultiContext(n_gpus=0, eager_reads=False) as mctx:
        g = mctx.check_linearizable(h, m)
    with native.Context(eager_reads=False) as ctx:
        s = [ctx.check_linearizable(p, m) for p in parts]
    assert g["valid"] == H.INVALID and g["n_failures"] == 1
    assert [x["valid"] for x in g["shards"]] == [s[0]["valid"], s[1]["valid"]] == [H.INVALID, H.VALID]
    assert g["shards"][0]["witness_index"] == s[0]["shards"][0]["witness_index"]
