"""The closed form behind the exact empty-space skip (csrc/vt_device.cuh, advance_round), restated in numpy integer algebra and
checked on the CPU against the literal loop of dda.h:51-53 (`while (d <= tau && k < nmax) d = fl(d + e)`): same count, same bits.
The device function itself is checked against the literal loop by tests/test_gpu_configs.py::test_advance_until_equals_literal_loop."""
import numpy as np

f32 = np.float32


def literal(d, e, tau, nmax):
    d = d.copy(); k = np.zeros(d.shape, np.int32)
    for _ in range(int(nmax.max())):
        c = (d <= tau) & (k < nmax)
        if not c.any():
            break
        d = np.where(c, (d + e).astype(f32), d); k += c
    return d, k


def advance_binade(d, k, e, tau, nmax, quotient_error=1.0):
    """advance_binade of csrc/vt_device.cuh, line by line; `quotient_error` perturbs the approximate reciprocal the device uses."""
    def real_add(d, k):
        c = (d <= tau) & (k < nmax)
        return np.where(c, (d + e).astype(f32), d).astype(f32), k + c.astype(np.int32), c
    d, k, _ = real_add(d, k)
    d, k, _ = real_add(d, k)
    b2 = d.view(np.int32).astype(np.int64)
    d, k, third = real_add(d, k)
    b3 = d.view(np.int32).astype(np.int64)
    inc = b3 - b2
    N = tau.view(np.int32).astype(np.int64) - b3
    ic = np.maximum(inc, 1)
    qd = (N.astype(f32) * (f32(1.0) / ic.astype(f32) * f32(quotient_error)).astype(f32)).astype(f32).astype(np.int64)
    rem = N - qd * ic
    qd = np.where(rem < 0, qd - 1, np.where(rem >= ic, qd + 1, qd))
    J = np.minimum(qd, nmax - k)
    jump = third & (d <= tau) & (inc > 0) & (J > 0)
    d = np.where(jump, (b3 + J * inc).astype(np.int32).view(f32), d).astype(f32)
    k = k + np.where(jump, J, 0).astype(np.int32)
    d, k, _ = real_add(d, k)
    return d, k


def binade_end(t):
    return (t.view(np.int32) | np.int32(0x7fffff)).view(f32)


def closed_form(d, e, tau, nmax, quotient_error=1.0):
    """advance_until of csrc/vt_device.cuh (the test hook's single-axis form): binade after binade."""
    d = d.copy(); k = np.zeros(d.shape, np.int32); rounds = np.zeros(d.shape, np.int32)
    while True:
        cond = (d <= tau) & (k < nmax)
        if not cond.any():
            break
        rounds += cond
        tc = np.where(cond & (d.view(np.int32) >= 0), np.minimum(tau, binade_end(d)), f32(-1.0)).astype(f32)   # -1: no addition is made
        d2, k2 = advance_binade(d, k, e, tc, nmax, quotient_error)
        stuck = cond & (k2 == k)
        d = np.where(stuck, (d + e).astype(f32), d2).astype(f32)
        k = k2 + stuck.astype(np.int32)
    return d, k, rounds


def adversarial(seed, n):
    rng = np.random.RandomState(seed)
    e = (10.0 ** rng.uniform(-3, 5, n)).astype(f32)
    e[::11] = (e[::11].view(np.uint32) & np.uint32(0xFFFFF000)).view(f32)          # low bits zero: ties in many binades
    e[::13] = (e[::13].view(np.uint32) & np.uint32(0xFF800000)).view(f32)          # powers of two
    e[::19] = (e[::19].view(np.uint32) | np.uint32(1)).view(f32)                   # odd significand: a tie one binade up
    K = (10.0 ** rng.uniform(0, 3.3, n)).astype(f32)
    d = ((rng.uniform(0, 1, n).astype(f32) + np.floor(K)) * e).astype(f32)
    d[::7] = 0.0
    d[::17] = (e[::17] * f32(2.0 ** 20)).astype(f32)                               # d >> e
    d[::29] = (e[::29] * f32(2.0 ** 24)).astype(f32)                               # e below half an ulp of d: d never moves
    d[::31] = (e[::31] * f32(2.0 ** -5)).astype(f32)                               # e above d's binade
    nmax = rng.randint(1, 300, n).astype(np.int32)
    steps = rng.randint(0, 320, n).astype(f32)
    tau = ((d + steps * e) * rng.choice(f32([0.999, 1.0, 0.5, 1.001]), n)).astype(f32)
    ok = (e > 0) & (tau > 0) & np.isfinite(tau)
    return d[ok], e[ok], tau[ok], nmax[ok]


def test_closed_form_equals_literal_loop_adversarial():
    for seed, err in ((0, 1.0), (1, 1.0 + 2.0 ** -22), (2, 1.0 - 2.0 ** -22)):
        d, e, tau, nmax = adversarial(seed, 120000)
        a_d, a_k = literal(d, e, tau, nmax)
        b_d, b_k, _ = closed_form(d, e, tau, nmax, err)
        assert np.array_equal(a_k, b_k)
        assert np.array_equal(a_d.view(np.uint32), b_d.view(np.uint32))
        assert (a_k > 50).sum() > 10000 and (a_k == nmax).sum() > 1000 and (a_k == 0).sum() > 100


def test_closed_form_on_dda_operands_needs_few_rounds():
    """Operands as the DDA produces them (e = |1/direction|, d = t of the next crossing, runs of 8..128 voxels): exact, and one
    round (one jump) for most runs, two when the run crosses a binade."""
    rng = np.random.RandomState(3); n = 200000
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
    e = (1.0 / np.maximum(np.abs(v[:, 0]), 1e-5)).astype(f32)
    t = rng.uniform(0, 900, n).astype(f32)
    d = (t + rng.uniform(0, 1, n).astype(f32) * e).astype(f32)
    nmax = rng.randint(3, 129, n).astype(np.int32)
    tau = (t + rng.uniform(8, 128, n)).astype(f32)
    a_d, a_k = literal(d, e, tau, nmax)
    b_d, b_k, rounds = closed_form(d, e, tau, nmax)
    assert np.array_equal(a_k, b_k) and np.array_equal(a_d.view(np.uint32), b_d.view(np.uint32))
    assert (rounds <= 2).mean() > 0.9 and a_k.mean() > 15


def test_three_axes_stop_at_a_common_threshold():
    """dda_skip's pass structure: all three axes advance to the end of the binade of the smallest dis (or to tau); the state must be
    the one the merged literal iteration of dda.h:51-53 reaches when it has processed every value <= that threshold."""
    rng = np.random.RandomState(9); n = 60000
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1)[:, None]
    e = (1.0 / np.maximum(np.abs(v), 1e-5)).astype(f32)
    t0 = (10.0 ** rng.uniform(-1, 3, n)).astype(f32)
    d = (t0[:, None] + rng.uniform(0, 1, (n, 3)).astype(f32) * e).astype(f32)
    tau = (t0 + rng.uniform(4, 200, n)).astype(f32)
    nmax = np.full(n, 1 << 20, np.int32)
    # literal merged loop: step every axis whose dis equals the minimum, while the minimum is <= tau
    ld = d.copy(); lk = np.zeros((n, 3), np.int32)
    for _ in range(2000):
        m = ld.min(axis=1)
        go = m <= tau
        if not go.any():
            break
        mask = (ld == m[:, None]) & go[:, None]
        ld = np.where(mask, (ld + e).astype(f32), ld); lk += mask
    assert not (ld.min(axis=1) <= tau).any()
    cd = d.copy(); ck = np.zeros((n, 3), np.int32)
    for _ in range(40):                                       # passes: binade after binade until tau
        tmin = cd.min(axis=1)
        tc = np.minimum(tau, binade_end(tmin)).astype(f32)
        live = tc > tmin
        if not live.any():
            break
        tc = np.where(live, tc, f32(-1.0)).astype(f32)
        for a in range(3):
            cd[:, a], ck[:, a] = advance_binade(cd[:, a].copy(), ck[:, a].copy(), e[:, a].copy(), tc, nmax)
    assert np.array_equal(ck, lk) and np.array_equal(cd.view(np.uint32), ld.view(np.uint32))
