"""Hand known-answer histories in a compact notation.

`"0:inv write 1, 0:ok write 1, 1:inv read, 1:ok read nil"` -> list of Jepsen-style op maps.
Values: ints, nil, [old new] (cas), #{1 2} (set read), {1 -3 2 3} (bank read), t(1 2 3) = transfer
debit 1 credit 2 amount 3.
"""
import re

from jepsen_tigerbeetle_b200.history import INVALID, UNKNOWN, VALID  # noqa: F401


def _value(f, txt):
    txt = txt.strip()
    if txt == "" or txt == "nil":
        return None
    if txt.startswith("#{"):
        return set(int(x) for x in txt[2:-1].split())
    if txt.startswith("[") and f == "read":
        return [int(x) for x in txt[1:-1].split()]
    if txt.startswith("["):
        return [int(x) for x in txt[1:-1].split()]
    if txt.startswith("{"):
        xs = txt[1:-1].split()
        return {int(xs[i]): (None if xs[i + 1] == "nil" else int(xs[i + 1])) for i in range(0, len(xs), 2)}
    if txt.startswith("t("):
        d, c, a = (int(x) for x in txt[2:-1].split())
        return {"debit-acct": d, "credit-acct": c, "amount": a}
    return int(txt)


def ops(text, times=None):
    out = []
    for i, item in enumerate(x.strip() for x in text.split(",")):
        m = re.match(r"(\d+):(inv|ok|fail|info)\s+(\w+)\s*(.*)", item)
        assert m, item
        p, ty, f, val = m.groups()
        ty = {"inv": "invoke"}.get(ty, ty)
        out.append({"process": int(p), "type": ty, "f": f, "value": _value(f, val), "index": i,
                    "time": (times[i] if times else i * 1000)})
    return out


# (name, model, history, expected verdict, expected witness index or None)
REGISTER_KATS = [
    ("B1", "cas-register", "0:inv write 1, 0:ok write 1, 0:inv read, 0:ok read 1", VALID, None),
    ("B2", "cas-register", "0:inv write 1, 0:ok write 1, 0:inv write 2, 0:ok write 2, 1:inv read, 1:ok read 1", INVALID, 5),
    ("B3", "cas-register", "0:inv write 1, 1:inv read, 1:ok read 1, 0:ok write 1", VALID, None),
    ("B4", "cas-register", "0:inv write 1, 0:ok write 1, 0:inv write 2, 1:inv read, 1:ok read 1, 0:ok write 2", VALID, None),
    ("B5", "cas-register", "0:inv write 0, 0:ok write 0, 1:inv write 1, 1:info write 1, 2:inv read, 2:ok read 1, 2:inv read, 2:ok read 0", INVALID, 7),
    ("B6", "cas-register", "0:inv write 0, 0:ok write 0, 1:inv write 1, 1:info write 1, 2:inv read, 2:ok read 1, 2:inv read, 2:ok read 1", VALID, None),
    ("B7", "cas-register", "0:inv write 0, 0:ok write 0, 1:inv write 1, 1:info write 1, 2:inv read, 2:ok read 0, 2:inv read, 2:ok read 0", VALID, None),
    ("B8", "cas-register", "0:inv write 0, 0:ok write 0, 0:inv cas [0 1], 0:ok cas [0 1], 0:inv read, 0:ok read 1", VALID, None),
    ("B9", "cas-register", "0:inv write 2, 0:ok write 2, 0:inv cas [0 1], 0:ok cas [0 1]", INVALID, 3),
    ("B10", "cas-register", "0:inv write 2, 0:ok write 2, 0:inv cas [0 1], 0:fail cas [0 1], 0:inv read, 0:ok read 2", VALID, None),
    ("B11", "cas-register", "0:inv write 1, 0:ok write 1, 1:inv read, 1:ok read nil", VALID, None),
    ("B12", "cas-register", "0:inv write 1, 1:inv write 2, 0:ok write 1, 1:ok write 2, 2:inv read, 2:ok read 1, 3:inv read, 3:ok read 2", INVALID, 7),
    ("B13", "cas-register", "0:inv write 1, 1:inv write 2, 0:ok write 1, 1:ok write 2, 2:inv read, 3:inv read, 2:ok read 1, 3:ok read 2", INVALID, None),
    ("B14", "cas-register", "0:inv write 1, 1:inv write 2, 2:inv read, 3:inv read, 0:ok write 1, 1:ok write 2, 2:ok read 1, 3:ok read 2", VALID, None),
    # plain register model: same answers on cas-free histories, cas is inconsistent
    ("B2r", "register", "0:inv write 1, 0:ok write 1, 0:inv write 2, 0:ok write 2, 1:inv read, 1:ok read 1", INVALID, 5),
    ("B14r", "register", "0:inv write 1, 1:inv write 2, 2:inv read, 3:inv read, 0:ok write 1, 1:ok write 2, 2:ok read 1, 3:ok read 2", VALID, None),
    # a read of a value nobody wrote from the nil initial state
    ("R1", "cas-register", "0:inv read, 0:ok read 3", INVALID, 1),
    ("R2", "cas-register", "0:inv read, 0:ok read nil", VALID, None),
    # never-completed invoke behaves like :info
    ("R3", "cas-register", "0:inv write 5, 1:inv read, 1:ok read 5", VALID, None),
]

SET_KATS = [
    ("B20", "set", "0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{1}", VALID, None),
    ("B21", "set", "0:inv add 1, 0:ok add 1, 1:inv read, 1:ok read #{}", INVALID, 3),
    ("B22", "set", "0:inv add 1, 1:inv read, 1:ok read #{}, 0:ok add 1", VALID, None),
    ("B23", "set", "0:inv add 1, 0:info add 1, 1:inv read, 1:ok read #{1}, 1:inv read, 1:ok read #{}", INVALID, 5),
    ("B24", "set", "1:inv read, 1:ok read #{7}", INVALID, 1),
]

_Z = "3 0 4 0 5 0 6 0 7 0 8 0"
BANK_KATS = [
    ("B40", "bank", f"0:inv transfer t(1 2 3), 0:ok transfer t(1 2 3), 0:inv read, 0:ok read {{1 -3 2 3 {_Z}}}", VALID, None),
    ("B41", "bank", f"0:inv transfer t(1 2 3), 0:ok transfer t(1 2 3), 0:inv read, 0:ok read {{1 -3 2 0 {_Z}}}", INVALID, 3),
    ("B42", "bank", "0:inv transfer t(1 2 3), 0:ok transfer t(1 2 3), 0:inv transfer t(2 3 1), 0:ok transfer t(2 3 1), "
                    "1:inv read, 1:ok read {1 -3 2 3 3 0 4 0 5 0 6 0 7 0 8 0}", INVALID, 5),
    ("B43", "bank", "0:inv read, 0:ok read {1 0 2 0 3 0 4 0 5 0 6 0 7 0 8 0 9 0}", INVALID, 1),
    ("B46", "bank", "0:inv transfer t(1 2 3), 1:inv read, 1:ok read {1 0 2 0 3 0 4 0 5 0 6 0 7 0 8 0}, 0:ok transfer t(1 2 3), "
                    "1:inv read, 1:ok read {1 -3 2 3 3 0 4 0 5 0 6 0 7 0 8 0}", VALID, None),
    ("B47", "bank", "0:inv transfer t(1 2 3), 0:info transfer t(1 2 3), 1:inv read, 1:ok read {1 -3 2 3 3 0 4 0 5 0 6 0 7 0 8 0}, "
                    "1:inv read, 1:ok read {1 0 2 0 3 0 4 0 5 0 6 0 7 0 8 0}", INVALID, 5),
    ("B48", "bank", "0:inv read, 0:ok read {1 nil 2 0 3 0 4 0 5 0 6 0 7 0 8 0}", INVALID, 1),
]
ALL_LIN_KATS = REGISTER_KATS + SET_KATS + BANK_KATS
