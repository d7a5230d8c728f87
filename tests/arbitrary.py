"""Seeded generator of ARBITRARY small histories (random interleavings and return values; most are not
linearizable) — the GPU-tier twin of tests/test_oracle_hypothesis.py."""
import numpy as np

VALUES = [0, 1, 2]


def arbitrary_history(model, rng, max_events=14, n_proc=3):
    ops, open_op, idx, next_elem = [], {}, 0, 1
    for _ in range(int(rng.integers(2, max_events + 1))):
        p = int(rng.integers(0, n_proc))
        if p not in open_op:
            v = None
            if model == "set":
                f = ["add", "read"][int(rng.integers(0, 2))]
                if f == "add":
                    v, next_elem = next_elem, next_elem + 1
            elif model == "bank":
                f = ["transfer", "read"][int(rng.integers(0, 2))]
                if f == "transfer":
                    d = int(rng.integers(1, 4)); c = int(rng.integers(1, 3)); c = c if c < d else c + 1
                    v = {"debit-acct": d, "credit-acct": c, "amount": int(rng.integers(1, 3))}
            else:
                kinds = ["read", "write", "cas"] if model == "cas-register" else ["read", "write"]
                f = kinds[int(rng.integers(0, len(kinds)))]
                if f == "write":
                    v = VALUES[int(rng.integers(0, 3))]
                elif f == "cas":
                    v = [VALUES[int(rng.integers(0, 3))], VALUES[int(rng.integers(0, 3))]]
            op = {"process": p, "type": "invoke", "f": f, "value": v, "index": idx, "time": idx * 1000}
            open_op[p] = op
            ops.append(op)
        else:
            inv = open_op.pop(p)
            typ = ["ok", "ok", "ok", "info", "fail"][int(rng.integers(0, 5))]
            v = inv["value"]
            if inv["f"] == "read" and typ == "ok":
                if model == "set":
                    v = set(int(x) for x in rng.integers(1, max(2, next_elem + 1), size=int(rng.integers(0, 4))))
                elif model == "bank":
                    v = {a: int(rng.integers(-3, 4)) for a in (1, 2, 3)}
                else:
                    v = [0, 1, 2, None][int(rng.integers(0, 4))]
            if typ == "fail" and inv["f"] in ("write", "add", "transfer"):
                typ = "ok"
            ops.append({"process": p, "type": typ, "f": inv["f"], "value": v, "index": idx, "time": idx * 1000})
        idx += 1
    return ops
