"""The certified comparisons of the wavefront sweep (cspm_pm_wave.cuh) rest on one inequality: for n non-negative
addends, the reference's row-major chain S and ANY other summation order T (here: lane partials, warp butterfly, a
short serial tail) satisfy |S - T| <= eps * T with eps = 4 n u, u = 2^-53.  This test emulates both orders in
numpy on adversarial and random addend sets and checks the inequality with its margin, plus the decision logic built
on it (an interval test never contradicts the exact comparison; ties are reported as undecided, not decided)."""
import numpy as np
import pytest

U = 2.0 ** -53


def seq_sum(t):
    s = 0.0
    for v in t:  # row-major chain, one rounding per add (pre_ss_pc.cc:84-116)
        s = s + v
    return s


def tree_sum(t, warps=20, lanes=32):
    """The kernel's order: chunks of 32 taps dealt to the warps in contiguous runs, every lane adds its taps in chunk
    order, xor-butterfly over the lanes, then the per-warp results are added in warp order."""
    n = len(t)
    chunks = (n + lanes - 1) // lanes
    per = (chunks + warps - 1) // warps
    padded = np.zeros(chunks * lanes)
    padded[:n] = t
    grid = padded.reshape(chunks, lanes)
    total = 0.0
    for w in range(warps):
        c0, c1 = w * per, min(chunks, (w + 1) * per)
        if c0 >= c1:
            continue
        acc = np.zeros(lanes)
        for c in range(c0, c1):
            acc = acc + grid[c]
        o = lanes // 2
        while o:
            acc = acc + acc[np.arange(lanes) ^ o]
            o //= 2
        total = total + acc[0]
    return total


def cases(rng, n):
    yield rng.uniform(0, 3, n)                                   # ordinary costs
    yield np.exp(rng.uniform(-70, 3, n))                         # 30 orders of magnitude (exp(-L1/10) weights)
    yield np.full(n, 2.8 * 0.1)                                  # saturated window: identical addends
    v = np.zeros(n); v[rng.integers(0, n, 5)] = rng.uniform(1, 1e6, 5); yield v   # sparse spikes
    yield np.sort(np.exp(rng.uniform(-40, 5, n)))                # ascending: worst case for a serial chain
    yield np.sort(np.exp(rng.uniform(-40, 5, n)))[::-1].copy()   # descending
    yield np.zeros(n)


@pytest.mark.parametrize("n", [25, 121, 1225, 3675, 3969])
def test_row_major_chain_and_tree_sum_agree_within_eps(n):
    rng = np.random.default_rng(n)
    eps = 4.0 * n * U
    worst = 0.0
    for rep in range(3):
        for t in cases(rng, n):
            S, T = seq_sum(t), tree_sum(t)
            assert abs(S - T) <= eps * T, (n, S, T)
            if T > 0:
                worst = max(worst, abs(S - T) / T)
    assert worst <= eps / 4  # the a-priori bound has a comfortable margin over what summation really does


def decide(T0, on0, T1, on1, m, eps):
    """The kernel's interval logic; returns best in {-1, 0, 1} or None when undecided."""
    lo0, hi0, lo1, hi1 = T0 * (1 - eps), T0 * (1 + eps), T1 * (1 - eps), T1 * (1 + eps)
    best, mlo, mhi = -1, m, m
    if on0:
        if hi0 < m:
            best, mlo, mhi = 0, lo0, hi0
        elif not lo0 >= m:
            return None
    if on1:
        if hi1 < mlo:
            best = 1
        elif not lo1 >= mhi:
            return None
    return best


def reference_decision(S0, on0, S1, on1, m):
    best, mn = -1, m
    if on0 and S0 < mn:
        best, mn = 0, S0
    if on1 and S1 < mn:
        best, mn = 1, S1
    return best


def test_interval_decisions_never_contradict_the_exact_ones():
    rng = np.random.default_rng(7)
    n = 1225
    eps = 4.0 * n * U
    undecided = decided = 0
    for rep in range(300):
        base = np.exp(rng.uniform(-30, 1, n))
        # candidates that are near-copies of the own plane's addends (ties and near-ties), and unrelated ones
        own = base
        c0 = base * (1 + rng.choice([0.0, 1e-16, -1e-16, 1e-3, -1e-3], n)) if rep % 2 else np.exp(rng.uniform(-30, 1, n))
        c1 = c0.copy() if rep % 3 == 0 else base * rng.uniform(0.9, 1.1)
        c0, c1 = np.abs(c0), np.abs(c1)
        m, S0, S1 = seq_sum(own), seq_sum(c0), seq_sum(c1)
        T0, T1 = tree_sum(c0), tree_sum(c1)
        for on0, on1 in ((True, True), (True, False), (False, True)):
            got = decide(T0, on0, T1, on1, m, eps)
            if got is None:
                undecided += 1
            else:
                decided += 1
                assert got == reference_decision(S0, on0, S1, on1, m), (rep, on0, on1, T0, T1, m, S0, S1)
    assert decided > 100 and undecided > 10  # both branches exercised


def test_partial_slot_bound_of_the_chunk_partition():
    """Mirror of the kernel's chunk partition (k_spatial_wave: total = E * cpt chunks, per = ceil(total / warps),
    warp w owns [w * per, (w + 1) * per)): the number of candidate evaluations a warp touches never exceeds
    kWvSlots = ceil(2 * kWvTasks / warps) + 1, and the decision phase finds every partial where it expects it."""
    tasks = 32
    for warps in (16, 20, 24, 32):
        slots = (2 * tasks + warps - 1) // warps + 1
        for cpt in (1, 4, 39, 117, 125, 1000):
            for E in range(1, 2 * tasks + 1):
                total = E * cpt
                per = max(1, (total + warps - 1) // warps)
                covered = [set() for _ in range(E)]
                for w in range(warps):
                    g0, g1 = w * per, min(total, (w + 1) * per)
                    if g0 >= g1:
                        continue
                    first, last = g0 // cpt, (g1 - 1) // cpt
                    assert last - first + 1 <= slots, (warps, cpt, E, w)
                    for e in range(first, last + 1):
                        covered[e].add(w)
                for e in range(E):  # decision phase: warps w_first .. w_last contribute, slot = e - first_eval(w)
                    w_first, w_last = (e * cpt) // per, ((e + 1) * cpt - 1) // per
                    assert covered[e] == set(range(w_first, w_last + 1)), (warps, cpt, E, e)
                    for w in range(w_first, w_last + 1):
                        assert 0 <= e - (w * per) // cpt < slots
