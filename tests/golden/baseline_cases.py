"""The BASELINE.json-size parity cases, shared by make_baseline_golden.py (oracle to completion ->
baseline_golden.json) and tests/test_gpu_baseline_configs.py (GPU through the C ABI vs those records)."""
from jepsen_tigerbeetle_b200 import history as H, synth

MODEL = {"register": H.MODEL_REGISTER, "cas-register": H.MODEL_CAS_REGISTER, "set": H.MODEL_SET, "bank": H.MODEL_BANK}


def model_of(name):
    return H.make_model(MODEL[name], accounts=range(1, 9)) if name == "bank" else H.make_model(MODEL[name])


CASES = {}
# C3: the exact bench.py instance, seeds 1-3, valid + one stale read, Knossos-exact and eager-read spaces
for seed in (1, 2, 3):
    for stale in (False, True):
        for eager in (False, True):
            CASES[f"c3_bank10k32_think5_seed{seed}_{'stale' if stale else 'valid'}_{'eager' if eager else 'exact'}"] = dict(
                kind="lin", model="bank", gen="c3", seed=seed, stale=stale, eager=eager, max_configs=450_000_000)
# C3 EXACTLY as SURVEY 8(d) writes it (tau_think 0: 6-7e9 configurations in the Knossos-exact space): no exhaustive CPU
# search can finish these, so the fixture comes from the oracle's reduced bank decider (ALGO_LAZY_BANK, ~1e5
# configurations, equal to knossos.wgl on every history both can finish) — verdict, witness, previous-ok; its configuration
# count is NOT the exhaustive count and is not compared
for seed in (1, 2, 3):
    for stale in (False, True):
        CASES[f"c3_bank10k32_think0_seed{seed}_{'stale' if stale else 'valid'}_exact"] = dict(
            kind="lin", model="bank", gen="c3", seed=seed, stale=stale, eager=False, think_ms=0.0, oracle_algo=5, compare_counts=False)
# C2 at full size is already a pytest case run against the live oracle (test_config_c2)
# C4: 100k-op set-full, 64 clients, K = 64 and K = 8 ledgers; clean and one poisoned ledger
for K in (64, 8):
    for poisoned in (False, True):
        tag = f"c4_setfull100k_K{K}_{'poisoned' if poisoned else 'clean'}"
        CASES[tag + "_scan"] = dict(kind="setfull", gen="c4", K=K, poisoned=poisoned)
    CASES[f"c4_setfull100k_K{K}_clean_wglset"] = dict(kind="lin", model="set", gen="c4", K=K, poisoned=False, eager=True,
                                                      threads=8, max_configs=50_000_000)
# C5: 50k-op cas-register, 30 % :info, K = 256 keys (clean + the poisoned key bench.py uses), K = 8 monster (budget)
for poisoned in (False, True):
    CASES[f"c5_cas50k_K256_pinfo30_{'poisoned' if poisoned else 'clean'}"] = dict(
        kind="lin", model="cas-register", gen="c5", K=256, poisoned=poisoned, eager=True, threads=8, max_configs=50_000_000)
CASES["c5_cas50k_K8_monster"] = dict(kind="lin", model="cas-register", gen="c5", K=8, poisoned=False, eager=True,
                                     threads=8, max_configs=50_000_000)


def build_history(case):
    if case["gen"] == "c3":
        return synth.generate(synth.SynthSpec("bank", 10000, 32, case["seed"], tau_think_ns=case.get("think_ms", 5.0) * 1e6,
                                              stale_read=case["stale"]))
    if case["gen"] == "c4":
        h = synth.config_c4(seed=1, n_keys=case["K"])
        return synth.poison_c4(h, 5) if case["poisoned"] else h
    if case["gen"] == "c5":
        h = synth.config_c5(seed=1, n_keys=case["K"])
        return synth.poison_c5(h, 7) if case["poisoned"] else h
    raise ValueError(case["gen"])
