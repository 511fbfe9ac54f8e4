"""CPU pins of the device-side logic of the two-records-per-trip blend backward (csrc/blend_stream.cu, GSB_BWD_PAIR;
DESIGN.md section 5.2) that no host code sees.  Pure numpy float32.

1. Validity folded into the operands: an invalid pixel continues with alpha = 0 and a pixel outside the gate with
   opacity * vis = 0.  The per-pixel recurrence of rasterize_to_pixels_bwd.cu:192-241 written that way -- no select
   after the validity decision -- must give bit-for-bit the sums of the form with selects.
2. The halving butterfly over 16 values (`reduce_spread<16>`, csrc/blend.cuh) and the lanes that publish it: every one
   of the 2 x 8 values of a pair of records is published exactly once, by a lane whose id select ((lane & 16) ? second
   : first) picks the record the value belongs to, at offset slot & 7.
"""
import numpy as np

F = np.float32
ALPHA_MAX, ALPHA_MIN = F(0.999), F(1.0 / 255.0)


def _records(rng, n):
    return dict(
        sigma=rng.uniform(-0.2, 6.0, n).astype(F),       # some negative: the reference skips those
        opac=rng.uniform(0.0, 1.0, n).astype(F) * (rng.uniform(size=n) > 0.1),  # some exactly 0
        boost=np.where(rng.uniform(size=n) > 0.9, F(40.0), F(1.0)).astype(F),   # some clamped at 0.999
        c=rng.uniform(0, 1, (n, 3)).astype(F),
        idx=np.arange(n),
    )


def _walk(rec, v, v_a, t_final, last, folded):
    """Back-to-front walk of one pixel; returns per record (fac * v_k sums' common factor fac, -v_sigma)."""
    n = len(rec["sigma"])
    T, B = F(t_final), F(0)
    out = np.zeros((n, 2), F)
    for i in range(n - 1, -1, -1):
        vis = np.exp(-rec["sigma"][i]).astype(F)
        ov = F(rec["opac"][i] * rec["boost"][i] * vis)
        alpha = min(ALPHA_MAX, ov)
        valid = i <= last and rec["sigma"][i] >= 0 and alpha >= ALPHA_MIN
        gate = valid and ov <= ALPHA_MAX
        a2 = F(rec["c"][i] @ v)
        if folded:
            a_eff = alpha if valid else F(0)
            g_ov = ov if gate else F(0)
            nra = F(1) / F(a_eff - F(1))          # -1 exactly for an invalid pixel
            ra = -nra
            Tn = F(T * ra)
            T = Tn                                # no select: T * 1 = T
            fac = F(a_eff * Tn)                   # 0 * T = 0
            v_alpha = F(F(Tn * a2) + F(nra * B)) + F(F(t_final * ra) * v_a)
            B = F(F(fac * a2) + B)
            nvs = F(g_ov * v_alpha)
        else:
            nra = F(1) / F(alpha - F(1))
            ra = -nra
            Tn = F(T * ra)
            fac = F(alpha * Tn) if valid else F(0)
            v_alpha = F(F(Tn * a2) + F(nra * B)) + F(F(t_final * ra) * v_a)
            if valid:
                T = Tn
            B = F(F(fac * a2) + B)
            nvs = F(ov * v_alpha) if gate else F(0)
        out[i] = (fac, nvs)
    return out


def test_validity_folded_into_the_operands_is_bit_identical():
    rng = np.random.default_rng(1)
    for trial in range(20):
        n = 120
        rec = _records(rng, n)
        v = rng.normal(size=3).astype(F)
        v_a = F(rng.normal())
        last = int(rng.integers(n // 2, n))  # records beyond the pixel's last contributor are invalid
        t_final = F(rng.uniform(1e-3, 0.5))
        want = _walk(rec, v, v_a, t_final, last, folded=False)
        got = _walk(rec, v, v_a, t_final, last, folded=True)
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32)) or np.array_equal(want, got)
        assert np.isfinite(got).all()
        assert (got[:, 0] != 0).sum() > 5  # the walk is not vacuous


def _spread_slot(nv, lane):
    slot, alive, dist = 0, nv, 16
    while alive > 1:
        slot += (alive >> 1) if (lane & dist) else 0
        alive >>= 1
        dist >>= 1
    return slot


def _reduce_spread(vals):
    """vals[lane][k] -> v[lane][0] after the halving butterfly of reduce_spread<NV> (csrc/blend.cuh)."""
    v = vals.copy()
    nv = v.shape[1]
    alive, dist = nv, 16
    while alive > 1:
        half = alive >> 1
        nxt = v.copy()
        for lane in range(32):
            upper = bool(lane & dist)
            for q in range(half):
                keep = v[lane][half + q] if upper else v[lane][q]
                # the partner sends ITS `send`: an upper partner sends v[q], a lower partner v[half + q]
                partner_upper = bool((lane ^ dist) & dist)
                send_partner = v[lane ^ dist][q] if partner_upper else v[lane ^ dist][half + q]
                nxt[lane][q] = keep + send_partner
        v = nxt
        alive >>= 1
        dist >>= 1
    while dist > 0:
        nxt = v.copy()
        for lane in range(32):
            nxt[lane][0] = v[lane][0] + v[lane ^ dist][0]
        v = nxt
        dist >>= 1
    return v[:, 0]


def test_pair_butterfly_publishes_every_value_once_to_the_right_record():
    rng = np.random.default_rng(2)
    vals = rng.integers(-50, 50, (32, 16)).astype(np.float64)  # integers: sums are exact in any order
    # values 0..7 belong to the first record of the pair, 8..15 to the second
    totals = vals.sum(axis=0)
    reduced = _reduce_spread(vals)
    published = {}
    for lane in range(32):
        slot16 = _spread_slot(16, lane)
        assert (slot16 >= 8) == bool(lane & 16)  # one id select serves the main group and the tail
        if lane & 1:
            continue  # pub16
        record, off = (1 if lane & 16 else 0), slot16 & 7
        assert (record, off) not in published
        published[(record, off)] = reduced[lane]
        assert reduced[lane] == totals[slot16]
    assert len(published) == 16
    for (record, off), val in published.items():
        assert val == totals[record * 8 + off]
    # the two `sum w` of the pair: reduce_spread<2>, published by lanes 0 and 16
    tail = rng.integers(-50, 50, (32, 2)).astype(np.float64)
    red2 = _reduce_spread(tail)
    assert red2[0] == tail[:, 0].sum() and red2[16] == tail[:, 1].sum()
    assert _spread_slot(2, 0) == 0 and _spread_slot(2, 16) == 1


def _forward_pixel(alphas, sigmas, colours, words, group):
    """One pixel of the forward consumer loop (blend_stream_fwd_kernel): records taken `group` at a time, a missing record
    of the last trip standing in with alpha = 0.  T <= 0 marks a finished pixel (-T is its transmittance)."""
    T, out, cur = F(1), np.zeros(3, F), 0
    n = len(alphas)
    for base in range(0, n, group):
        trip = []
        for r in range(group):
            i = base + r
            have = i < n
            j = i if have else base  # (the kernel re-reads the trip's first record)
            al = min(ALPHA_MAX, alphas[j]) if have else F(0)
            trip.append((sigmas[j], al, colours[j], words[j]))
        for sigma, al, c, wt in trip:
            next_T = F(T * F(F(al * F(-1)) + F(1)))
            ok = not (sigma < 0 or al < ALPHA_MIN)
            take = ok and not (next_T <= F(1e-4))
            vis = F(al * T) if take else F(0)
            out = (c * vis + out).astype(F)
            T = next_T if take else (-abs(T) if ok else T)
            cur = wt if take else cur
    return T, out, cur


def test_forward_walk_in_groups_equals_the_record_by_record_walk():
    rng = np.random.default_rng(3)
    for trial in range(30):
        n = int(rng.integers(1, 90))
        sig = rng.uniform(-0.1, 5.0, n).astype(F)
        al = (rng.uniform(0, 1, n) * np.exp(-sig)).astype(F) * np.where(rng.uniform(size=n) > 0.8, F(30), F(1)).astype(F)
        col = rng.uniform(0, 1, (n, 3)).astype(F)
        words = [(i << 4) | 0xF for i in range(1, n + 1)]  # never 0
        ref = _forward_pixel(al, sig, col, words, 1)
        for group in (2, 3, 4):
            got = _forward_pixel(al, sig, col, words, group)
            assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and got[2] == ref[2], (trial, group)
    # a saturating pixel stops taking and keeps its transmittance negated
    T, out, cur = _forward_pixel(np.full(40, 0.9, F), np.zeros(40, F), np.ones((40, 3), F), list(range(16, 16 * 41, 16)), 2)
    assert T < 0 and 0 < -T <= 1e-3 and cur == 16 * 4
