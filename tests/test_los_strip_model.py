"""CPU model of the strip line-of-sight kernel (csrc/trrt_los.cuh) against the oracle's literal search.lineofsight
(search.py:35-94): the strip layout, the 8-slot block step with its virtual prefix, and the closed-form jump the
cooperative tail uses.  Pure numpy; the GPU kernel itself is compared with the oracle in test_gpu_parity.py."""
import numpy as np
import pytest

from tests import util


def build_strips(free):
    """[orientation][K][c][16] bytes: byte j bit i = pixel (8c+i, 8K-8+j) (o = 0) or (8K-8+j, 8c+i) (o = 1)."""
    n = free.shape[0]
    tp = (n + 7) // 8
    pad = np.zeros((8 * tp + 16, 8 * tp + 16), bool)  # image at offset (8, 8), indexed [y + 8, x + 8]
    pad[8:8 + n, 8:8 + n] = free
    K, c, j, i = np.meshgrid(np.arange(tp + 1), np.arange(tp), np.arange(16), np.arange(8), indexing="ij")
    o0 = pad[8 * K + j, 8 + 8 * c + i]
    o1 = pad[8 + 8 * c + i, 8 * K + j]
    return np.stack([np.packbits(o, axis=-1, bitorder="little")[..., 0] for o in (o0, o1)]), tp


F32 = np.float32
MAGIC, MAGIC_BITS = F32(12582912.0), 0x4B400000


def fma(a, b, c):
    """fp32 fused multiply-add: the product of two fp32 numbers plus an fp32 is exact in fp64 here (|values| < 2^24)."""
    return F32(np.float64(a) * np.float64(b) + np.float64(c))


def ray_consts(dmaj2, dmin2):
    rcp = F32(1.0) / F32(dmaj2)  # correctly rounded, like __frcp_rn
    return dict(dmaj2=dmaj2, dmin2=dmin2, rcp=rcp, slope=F32(F32(dmin2) * rcp), hrm=F32(F32(F32(0.5) * rcp) - F32(0.5)))


def floor_div_fma(x, c):
    """floor(x / dmaj2) the way the kernel computes it (trrt_los.cuh, TRRT_LOS_MAGIC)."""
    return int(F32(fma(F32(x), c["rcp"], c["hrm"]) + MAGIC).view(np.uint32)) - MAGIC_BITS


def block_step(strips, tp, orient, A, aend, klo, neg, c, b, u):
    """One block of one ray, as strip_block() does it; returns (blocked, b', u')."""
    entry = strips[orient, (b >> 3) + (0 if neg else 1), A >> 3]
    sb = (b & 7) + (1 if neg else 0)
    base = fma(F32(u), c["rcp"], c["hrm"])
    rows = [0] + [int(F32(fma(c["slope"], F32(k), base) + MAGIC).view(np.uint32)) - MAGIC_BITS for k in range(1, 9)]
    assert rows == [(u + k * c["dmin2"]) // c["dmaj2"] for k in range(9)]
    window = entry[sb:sb + 8]
    khi = min(7, aend - A)
    blocked = False
    for k in range(klo, khi + 1):
        byte = window[7 - rows[k]] if neg else window[rows[k]]
        blocked |= not (byte >> k) & 1
    j = rows[8]
    return blocked, b + (-j if neg else j), u + 8 * c["dmin2"] - j * c["dmaj2"]


def los_model(strips, tp, side, seg, coop_every=0):
    x0, y0, x1, y1 = (int(v) for v in seg)
    if not all(0 <= v < side for v in (x0, y0, x1, y1)):
        return False
    low = abs(y1 - y0) < abs(x1 - x0)
    p0, q0, p1, q1 = (x0, y0, x1, y1) if low else (y0, x0, y1, x1)
    if p0 > p1:
        p0, q0, p1, q1 = p1, q1, p0, q0
    dmaj, dq = p1 - p0, q1 - q0
    neg, dmaj2, dmin2 = dq < 0, (2 * dmaj if dmaj else 2), 2 * abs(dq)
    c = ray_consts(dmaj2, dmin2)
    klo, a = p0 & 7, p0 & ~7
    # the literal recurrence of search.py:58-94 from (q0, D0), to compare the state at every block start with
    bl, Dl = q0, dmin2 - dmaj
    w = (dmaj - 1 if dmaj else 0) - klo * dmin2 + 7 * dmaj2
    qd = floor_div_fma(w, c)
    assert qd == w // dmaj2
    u, b = w - qd * dmaj2, q0 + ((7 - qd) if neg else (qd - 7))
    first = True
    while a <= p1:
        if not first and dmaj:
            assert (b, u) == (bl, Dl - (dmin2 - 2 * dmaj + 1)), "phase / coordinate at a block start"
            if coop_every:  # the cooperative tail: m blocks ahead in one go
                m = coop_every
                bj, Dj = bl, Dl
                for _ in range(8 * m):
                    if Dj > 0:
                        bj += -1 if neg else 1
                        Dj -= dmaj2
                    Dj += dmin2
                U = u + m * 8 * dmin2
                steps = U // dmaj2
                assert (U - steps * dmaj2, b + (-steps if neg else steps)) == (Dj - (dmin2 - 2 * dmaj + 1), bj)
        blocked, b, u = block_step(strips, tp, 0 if low else 1, a, p1, klo, neg, c, b, u)
        for _ in range(8 - klo):  # advance the literal recurrence over the real slots of this block
            if Dl > 0:
                bl += -1 if neg else 1
                Dl -= 2 * dmaj
            Dl += dmin2
        if blocked:
            return False
        a += 8
        klo = 0
        first = False
    return True


def test_floor_div_fma_boundaries():
    """The FFMA + magic-number floor is exact where it is hardest: numerators at and next to multiples of the divisor,
    divisors up to 2 * 32767 (maps of side 32768), quotients up to 15."""
    rng = np.random.default_rng(0)
    ds = [2, 4, 6, 10, 14, 16, 510, 512, 8190, 8192, 16382, 32766, 65532, 65534] + [int(2 * v) for v in rng.integers(1, 32768, 300)]
    for d in ds:
        c = ray_consts(d, 0)
        xs = set()
        for q in range(0, 16):
            xs.update((q * d - 1, q * d, q * d + 1, q * d + d // 2))
        xs.update(int(v) for v in rng.integers(0, 16 * d, 50))
        for x in xs:
            if 0 <= x < 16 * d and x < (1 << 20):
                assert floor_div_fma(x, c) == x // d, (x, d)


@pytest.mark.parametrize("n,seed", [(37, 1), (64, 2), (100, 3)])
def test_strip_model_vs_oracle(n, seed):
    from oracle import c_oracle as O
    free = util.synthetic_map(n, 0.08, 2, seed)
    strips, tp = build_strips(free)
    rng = np.random.default_rng(seed)
    seg = rng.integers(-2, n + 2, size=(3000, 4)).astype(np.int32)
    seg[:600, 2:] = seg[:600, :2] + rng.integers(-9, 10, size=(600, 2))
    seg[600:650, 2:] = seg[600:650, :2]
    ref = O.lineofsight_batch(free, seg)
    got = np.array([los_model(strips, tp, n, s, coop_every=(k % 4)) for k, s in enumerate(seg)])
    assert np.array_equal(got, ref)
    assert ref.any() and not ref.all()


def test_strip_model_clear_map_all_directions():
    """On an empty map every in-bounds segment is visible: exercises window offsets for all slopes and both signs."""
    n = 48
    strips, tp = build_strips(np.ones((n, n), bool))
    for x1 in range(0, n, 5):
        for y1 in range(0, n, 3):
            assert los_model(strips, tp, n, (23, 17, x1, y1), coop_every=2)
            assert los_model(strips, tp, n, (x1, y1, 0, n - 1), coop_every=3)


def test_strip_model_property_random_maps():
    """Property test (hypothesis): any square map from 1 x 1 up, any density, any segments (also out of bounds, zero
    length, reversed): the strip model and the oracle's literal search.lineofsight agree, and the result does not depend
    on the direction the segment is given in (search.py:47-56 canonicalises)."""
    from hypothesis import given, settings, strategies as st
    from oracle import c_oracle as O

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 40), st.floats(0.0, 0.6), st.integers(0, 2 ** 31 - 1))
    def check(n, p, seed):
        rng = np.random.default_rng(seed)
        free = rng.random((n, n)) >= p
        strips, tp = build_strips(free)
        seg = rng.integers(-1, n + 1, size=(60, 4)).astype(np.int32)
        ref = O.lineofsight_batch(free, seg)
        got = np.array([los_model(strips, tp, n, s, coop_every=int(rng.integers(0, 4))) for s in seg])
        assert np.array_equal(got, ref)
        rev = np.array([los_model(strips, tp, n, s[[2, 3, 0, 1]]) for s in seg])
        assert np.array_equal(rev, ref)

    check()
