"""-m gpu: the device primitives behind the path (exclusive scan, stable LSD radix sort), through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rdr():
    from houdini_gsplat_renderer_b200 import renderer as R
    r = R.GSplatRenderer(0)
    yield r
    r.close()


@pytest.mark.parametrize("n", [0, 1, 31, 4095, 4096, 4097, 70001, 1 << 20, 3_000_017])
def test_exclusive_scan(rdr, n):
    rng = np.random.default_rng(n)
    x = rng.integers(0, 50, n, dtype=np.uint32)
    got, total = rdr.exclusive_scan(x)
    want = np.concatenate([[0], np.cumsum(x, dtype=np.uint64)[:-1]]).astype(np.uint32) if n else x
    assert np.array_equal(got, want)
    assert total == int(x.sum(dtype=np.uint64))


def test_exclusive_scan_64bit_total(rdr):
    x = np.full(3_000_000, 2000, np.uint32)           # total 6e9 > 2^32
    _, total = rdr.exclusive_scan(x)
    assert total == 6_000_000_000


@pytest.mark.parametrize("n,lo,hi,kind", [
    (1, 0, 32, "rand"), (33, 0, 32, "rand"), (4096, 0, 32, "rand"), (4097, 0, 32, "rand"),
    (100_003, 0, 32, "rand"), (1_000_000, 0, 32, "float"), (1_000_000, 0, 32, "dups"),
    (300_000, 0, 13, "tiles"), (300_000, 0, 17, "tiles"), (50_000, 0, 1, "rand"), (2_500_000, 0, 32, "float"),
    (200_000, 0, 32, "same"),
    (400_000, 0, 17, "tiles"), (1_000_000, 0, 25, "rand"), (700_001, 0, 27, "rand"), (9000, 0, 9, "rand"),   # 9-bit digits
    (6_000_000, 0, 26, "rand"),
])
def test_radix_sort_is_stable(rdr, n, lo, hi, kind):
    rng = np.random.default_rng(n + hi)
    if kind == "rand":
        k = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        if hi < 32:
            k &= np.uint32((1 << hi) - 1)
    elif kind == "float":       # depth keys: bits of non-negative floats in a narrow exponent range
        k = (rng.random(n, dtype=np.float32) * 20 + 1.5).view(np.uint32)
    elif kind == "dups":
        k = rng.integers(0, 1000, n, dtype=np.uint64).astype(np.uint32) * np.uint32(65537)
    elif kind == "same":
        k = np.full(n, 0x3F800000, np.uint32)
    else:                       # tile ids in depth order: high bits hold junk the sort must ignore
        k = rng.integers(0, 1 << hi, n, dtype=np.uint64).astype(np.uint32) | np.uint32(0xA5000000)
    v = np.arange(n, dtype=np.uint32)[::-1].copy()
    ko, vo = rdr.sort_pairs(k, v, lo, hi)
    mask = np.uint64((1 << (hi - lo)) - 1)
    digit = (k.astype(np.uint64) >> np.uint64(lo)) & mask
    perm = np.argsort(digit, kind="stable")
    assert np.array_equal(ko, k[perm])
    assert np.array_equal(vo, v[perm])
