"""CPU proof-by-enumeration of the arithmetic identities the packed GRD layout and the folded gradient
term rely on (crossscalepatchmatch_b200/csrc/cspm_device.cuh "Packed GRD pixel", raw_cost_cg).
numpy float32 / float64 arithmetic is IEEE round-to-nearest without FMA contraction, like the kernels'
__f*_rn / __d*_rn operations, so bitwise equality here is bitwise equality there."""
import numpy as np


def _gray_f32_all_colours():
    """gray = (R*0.299f + G*0.587f) + B*0.114f for all 2^24 8-bit colours (cc/grd_cc.cpp:70-73, OpenCV 2.4 order)."""
    v = np.arange(256, dtype=np.float32)
    rg = (v[:, None] * np.float32(0.299) + v[None, :] * np.float32(0.587)).astype(np.float32)  # [R][G]
    return (rg[:, :, None] + (v * np.float32(0.114))[None, None, :]).astype(np.float32)        # [R][G][B]


def test_every_float_gray_is_a_multiple_of_2_pow_minus_27():
    g = _gray_f32_all_colours().astype(np.float64)
    scaled = g * 2.0 ** 27
    assert np.array_equal(scaled, np.floor(scaled)), "float gray must be an integer multiple of 2^-27"
    assert g.min() >= 0.0 and g.max() <= 255.0
    # hence G = g[x+1] - g[x-1] (exact in f64) is a multiple of 2^-27 with |G| <= 255 (and the border constant 3.0 too),
    # and n = G * 2^27 + 2^37 is a non-negative integer below 2^38: the 38-bit field of the packed word
    assert 255.0 * 2.0 ** 27 + 2.0 ** 37 < 2.0 ** 38 and -255.0 * 2.0 ** 27 + 2.0 ** 37 >= 0


def _random_gradients(rng, n):
    g = _gray_f32_all_colours().reshape(-1).astype(np.float64)
    a, b = g[rng.integers(0, g.size, n)], g[rng.integers(0, g.size, n)]
    G = a - b
    G[:16] = [0.0, 3.0, -3.0, 255.0, -255.0, 2.0, -2.0, 2.0 ** -27, -(2.0 ** -27), 1.0, 1.5, 2.0 + 2.0 ** -27, 2.0 - 2.0 ** -27,
              float(np.float32(0.114)), 255.0 - 2.0 ** -16, 128.0]  # all multiples of 2^-27, incl. the clamp boundary
    return G


def test_packed_word_decodes_to_the_exact_gradient_difference():
    rng = np.random.default_rng(1)
    Gq, Gp = _random_gradients(rng, 200_000), _random_gradients(rng, 200_000)[::-1].copy()

    def pack_decode(G, colour):
        n = (G * 2.0 ** 27).astype(np.int64) + (1 << 37)
        assert ((n >= 0) & (n < (1 << 38))).all()
        word = n.astype(np.uint64) | (colour.astype(np.uint64) << np.uint64(40))
        hi = (word >> np.uint64(32)).astype(np.uint32)
        lo = (word & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        key = hi & np.uint32(0xFFFFFF00)                      # SAD operand: B,G,R in bytes 1..3
        d_hi = (hi & np.uint32(0xFF)) | np.uint32(0x43300000)  # PRMT: byte 0 under the exponent of 2^52
        bits = (d_hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
        return key, bits.view(np.float64)

    col_q = rng.integers(0, 1 << 24, Gq.size, dtype=np.uint32)
    col_p = rng.integers(0, 1 << 24, Gq.size, dtype=np.uint32)
    kq, dq = pack_decode(Gq, col_q)
    kp, dp = pack_decode(Gp, col_p)
    assert np.array_equal(kq >> np.uint32(8), col_q) and np.array_equal(kp >> np.uint32(8), col_p)
    assert np.array_equal(dq, 2.0 ** 52 + (Gq * 2.0 ** 27 + 2.0 ** 37))
    D = dq - dp                                               # one f64 subtraction in the kernel
    assert np.array_equal(D, (Gq - Gp) * 2.0 ** 27), "difference of the decoded doubles = (Gq - Gp) * 2^27 exactly"

    # gradient term of the raw cost: reference form vs the two kernel forms, bit for bit
    k = 1 - 0.1                                               # (1 - ALPHA) as the host evaluates it
    k27 = np.ldexp(k, -27)
    ref = k * np.minimum(np.abs(Gq - Gp), 2.0)                # cc/grd_cc.cpp:16-18
    scaled = k27 * np.minimum(np.abs(D), 2.0 ** 28)           # power-of-two scaling of both factors
    kd = k27 * D
    folded = np.abs(np.where(np.abs(kd) > k27 * 2.0 ** 28, k27 * 2.0 ** 28, kd))  # raw_cost_cg
    assert np.array_equal(ref.view(np.int64), scaled.view(np.int64))
    assert np.array_equal(ref.view(np.int64), folded.view(np.int64))


def test_byte_sad_of_shifted_keys_is_the_colour_l1_distance():
    rng = np.random.default_rng(2)
    a = rng.integers(0, 256, (100_000, 3), dtype=np.int64)
    b = rng.integers(0, 256, (100_000, 3), dtype=np.int64)
    l1 = np.abs(a - b).sum(1)

    def key(c):  # B | G<<8 | R<<16, shifted into bytes 1..3
        return ((c[:, 0] | (c[:, 1] << 8) | (c[:, 2] << 16)) << 8).astype(np.uint32)

    ka, kb = key(a), key(b)
    sad4 = sum(np.abs(((ka >> np.uint32(8 * i)) & np.uint32(255)).astype(np.int64) - ((kb >> np.uint32(8 * i)) & np.uint32(255)).astype(np.int64)) for i in range(4))
    assert np.array_equal(sad4, l1)


def test_plane_disparity_extremes_are_at_the_window_corners():
    """window_cost's all-valid fast path: q_disp = a*q_x + (b*q_y + c), evaluated with rounded products and
    sums (pre_ss_pc.cc:88,99), is monotone in q_x for fixed q_y and in q_y for fixed q_x, so over a window its
    minimum and maximum — and those of its truncation — are attained at the four corners."""
    rng = np.random.default_rng(3)
    for _ in range(400):
        a, b = rng.normal(scale=rng.choice([1e-3, 0.05, 0.7, 30.0]), size=2)
        c = rng.uniform(-50, 300)
        x0 = int(rng.integers(0, 1900)); y0 = int(rng.integers(0, 1060))
        xs = np.arange(x0, x0 + int(rng.integers(1, 36)), dtype=np.float64)
        ys = np.arange(y0, y0 + int(rng.integers(1, 36)), dtype=np.float64)
        q = (a * xs)[None, :] + ((b * ys) + c)[:, None]        # same operations, same order, no FMA
        corners = q[[0, 0, -1, -1], [0, -1, 0, -1]]
        assert q.min() == corners.min() and q.max() == corners.max()
        assert np.trunc(q).min() == np.trunc(corners).min() and np.trunc(q).max() == np.trunc(corners).max()
