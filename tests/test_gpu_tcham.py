"""Tensor-core Hamming tiles (csrc/tcham.cu: tcgen05.mma.kind::i8 over +-1-expanded descriptors, accumulators in
tensor memory): the raw distance matrix against numpy popcounts — every distance is an exact integer, so the bar is
bit-exact.  The batch matchers built on the same tiles are checked end to end by test_gpu_parity*.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import svo
    c = svo.Context(1241, 376, nfeatures=2000, max_batch=1, lanes=1, max_rows=5000)
    yield c
    c.close()


def popcount_matrix(a, b):
    x = np.unpackbits(a[:, None, :] ^ b[None, :, :], axis=2)
    return x.sum(axis=2).astype(np.int32)


@pytest.mark.parametrize("na,nb", [(1, 1), (128, 128), (129, 127), (300, 515), (1000, 257), (37, 2000)])
def test_hamming_matrix_is_exact(ctx, na, nb):
    rng = np.random.default_rng(na * 10007 + nb)
    a = rng.integers(0, 256, (na, 32), dtype=np.uint8); b = rng.integers(0, 256, (nb, 32), dtype=np.uint8)
    # extremes: identical rows (d = 0), complements (d = 256), single-bit differences in every bit position
    b[0] = a[0]
    if nb > 1 and na > 1:
        b[1] = ~a[1]
    for k in range(min(na, nb, 256)):
        if k >= 2:
            b[k] = a[k]
            b[k, (k % 256) // 8] ^= np.uint8(1 << (k % 8))
    got = ctx.hamming_matrix(a, b)
    want = popcount_matrix(a, b)
    assert (got == want).all(), np.argwhere(got != want)[:5]
    assert got[0, 0] == 0
    if nb > 1 and na > 1:
        assert got[1, 1] == 256
