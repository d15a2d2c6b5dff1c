"""GPU soak: the CUDA path through the C ABI against the oracle on randomised scenes
(tests/test_fuzz_parity.py generators, far beyond the suite's seeds).
usage (on a B200): PYTHONPATH=. python tools/soak_gpu.py <first_seed> <last_seed>"""
import sys

from oracle import oracle_ffi  # test infrastructure: this is a test tool
from pbrs_b200 import _ffi
from tests.test_fuzz_parity import _check

o, g = oracle_ffi.load(), _ffi.load()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    for ext in (False, True):
        try:
            _check(o, g, seed, ext=ext)
        except AssertionError as e:
            bad += 1
            print("FAIL", seed, ext, str(e)[:200], flush=True)
print("done", lo, hi, "failures", bad)
sys.exit(1 if bad else 0)
