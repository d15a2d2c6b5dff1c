"""GPU soak: the CUDA path through the C ABI against the oracle on randomised scenes
(tests/test_fuzz_parity.py generators, far beyond the suite's seeds; `zoo`: the material / texture /
environment generator of tools/soak_parity.py).
usage (on a B200): PYTHONPATH=. python tools/soak_gpu.py <first_seed> <last_seed> [zoo]"""
import importlib.util
import os
import sys

from oracle import oracle_ffi  # test infrastructure: this is a test tool
from pbrs_b200 import _ffi
from tests.test_fuzz_parity import _check

o, g = oracle_ffi.load(), _ffi.load()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
zoo = len(sys.argv) > 3 and sys.argv[3] == "zoo"
check_zoo = None
if zoo:
    # reuse the generator without running that tool's main loop
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "soak_parity.py")).read()
    ns = {"__name__": "soak_parity_lib", "__file__": os.path.join(os.path.dirname(os.path.abspath(__file__)), "soak_parity.py")}
    exec(compile(src.split("\no, h = oracle_ffi.load(), hs_load()")[0], "soak_parity.py", "exec"), ns)
    check_zoo = ns["check_zoo"]
bad = 0
for seed in range(lo, hi):
    for ext in ((None,) if zoo else (False, True)):
        try:
            if zoo:
                check_zoo(o, g, seed)
            else:
                _check(o, g, seed, ext=ext)
        except AssertionError as e:
            bad += 1
            print("FAIL", seed, ext, str(e)[:200], flush=True)
print("done", lo, hi, "failures", bad)
sys.exit(1 if bad else 0)
