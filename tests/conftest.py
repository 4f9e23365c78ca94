import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_SETS = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "vec_*.npz")))
PROFILE_OF = {"tufa420_sprot": "tufa420"}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


class GoldenSet:
    """A committed reference-generated vector set (tests/golden/make_golden.py)."""

    def __init__(self, name):
        from biobench2_b200.oprofile import OProfile
        from biobench2_b200.seqdb import PackedDB
        self.name = name
        self.vec = np.load(os.path.join(GOLDEN, "vec_%s.npz" % name))
        self.oprofile = OProfile.load(os.path.join(GOLDEN, "prof_%s.npz" % PROFILE_OF.get(name, name)))
        self.db = PackedDB(self.vec["residues"], self.vec["offsets"], self.vec["lengths"])


_cache = {}


def golden_set(name):
    if name not in _cache:
        _cache[name] = GoldenSet(name)
    return _cache[name]


@pytest.fixture(params=GOLDEN_SETS)
def gset(request):
    return golden_set(request.param)


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="session")
def gpu_ctx():
    from biobench2_b200 import engine
    ctx = engine.Context(0)
    yield ctx
    ctx.close()
