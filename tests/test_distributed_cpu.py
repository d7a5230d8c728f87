"""Multi-rank host logic on CPU: world_size 2 over gloo.  The per-rank check function is the oracle
here (there is no GPU in this tier); the product path passes native.Context.check_linearizable."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from jepsen_tigerbeetle_b200 import distributed, history as H, synth


def test_assign_shards_lpt_balanced():
    parts = distributed.assign_shards([9, 1, 8, 2, 7, 3], 2)
    assert sorted(sum(parts, [])) == list(range(6))
    loads = [sum([9, 1, 8, 2, 7, 3][s] for s in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 3
    assert distributed.assign_shards([1.0], 4) == [[0], [], [], []]


def test_select_shards_roundtrip():
    h = synth.generate(synth.SynthSpec("set", 600, 12, 3, n_keys=5, p_info=0.05))
    sub = h.select_shards([3, 1])
    sub.validate()
    assert list(sub.key_ids) == [int(h.key_ids[3]), int(h.key_ids[1])]
    a, b = sub.shard(0), h.shard(3)
    for name in ("type", "f", "process", "index", "time_ns", "a", "payload_len", "payload"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, poison, q):
    import torch.distributed as dist
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h = synth.generate(synth.SynthSpec("cas-register", 1200, 24, 7, n_keys=6, grouped_keys=True, p_info=0.05, n_values=30,
                                       stale_read=poison))
    m = H.make_model(H.MODEL_CAS_REGISTER)
    fn = lambda sub: oracle.check_linearizable(sub, m, 3)["shards"]  # noqa: E731
    r = distributed.check_sharded(h, fn, rank, world, distributed.torch_all_reduce_max())
    q.put((rank, r["valid"], r["shard_valid"].tolist(), r["shard_witness"].tolist(), r["mine"]))
    dist.destroy_process_group()


@pytest.mark.parametrize("poison", [False, True])
def test_two_ranks_gloo_match_single_process(oracle_mod, poison):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, poison, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    h = synth.generate(synth.SynthSpec("cas-register", 1200, 24, 7, n_keys=6, grouped_keys=True, p_info=0.05, n_values=30,
                                       stale_read=poison))
    ref = oracle_mod.check_linearizable(h, H.make_model(H.MODEL_CAS_REGISTER), 3)
    want_valid = [s["valid"] for s in ref["shards"]]
    want_wit = [s["witness_index"] for s in ref["shards"]]
    for rank, valid, sv, sw, mine in got:
        assert sv == want_valid and sw == want_wit and valid == ref["valid"]
    assert sorted(got[0][4] + got[1][4]) == list(range(6))  # every shard checked exactly once
    if poison:
        assert ref["valid"] == H.INVALID  # one poisoned shard flips the global verdict
