"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/oracle_common.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  PARITY UNPINNED: the reference ships no golden vectors for this path and its
algorithms live in un-vendored JVM dependencies that cannot run in this image."""
from .binding import (ALGO_BRUTE, ALGO_LAZY_BANK, ALGO_LEVEL, ALGO_LINEAR, ALGO_WGL, ALGO_WGL_COMPACT, build, check_bank_totals,
                      check_linearizable, check_set_full, final_configs)

__all__ = ["ALGO_BRUTE", "ALGO_LAZY_BANK", "ALGO_LEVEL", "ALGO_LINEAR", "ALGO_WGL", "ALGO_WGL_COMPACT", "build",
           "check_linearizable", "check_set_full", "check_bank_totals", "final_configs"]
