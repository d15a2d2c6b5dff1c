"""CPU soak: the product's stage functions (host build, tests/hostsim) against the oracle on
randomised scenes far beyond the seeds of tests/test_fuzz_parity.py.
usage: python tools/soak_parity.py <first_seed> <last_seed>
Every seed is run without and with the simple shapes, under path depth 4 (the suite's check), the
direct integrator, 4 spp, and path depth 6 without jitter.  Round 1: 1 400 + 1 440 renders, no
mismatch after the far-root flag."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_ffi  # noqa: E402  (test infrastructure: this is a test tool)
from pbrs_b200 import _capi as K  # noqa: E402
from tests.hostsim import load as hs_load  # noqa: E402
from tests.test_fuzz_parity import _check, random_scene  # noqa: E402
from tests.util import assert_radiance_close, assert_stats_close  # noqa: E402

o, h = oracle_ffi.load(), hs_load()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    for ext in (False, True):
        try:
            _check(o, h, seed, ext=ext)
            sd = random_scene(seed, ext=ext)
            ho, hp = sd.realize(o), sd.realize(h)
            for kw in (dict(integrator="direct", msaa=1, max_depth=5), dict(integrator="path", msaa=2, max_depth=2),
                       dict(integrator="path", msaa=1, max_depth=6, extra=K.FLAG_NO_JITTER)):
                fl = K.FLAG_COUNT_TRAVERSAL | kw.pop("extra", 0)
                fa, sa = ho.render_samples(flags=fl, **kw)
                fb, sb = hp.render_samples(flags=fl, **kw)
                assert_radiance_close(fb, fa, f"seed {seed} {kw}", outliers=2e-3)
                assert_stats_close(sb, sa, f"seed {seed} {kw}", rel=2e-3)
        except AssertionError as e:
            bad += 1
            print("FAIL", seed, ext, str(e)[:200], flush=True)
print("done", lo, hi, "failures", bad)
sys.exit(1 if bad else 0)
