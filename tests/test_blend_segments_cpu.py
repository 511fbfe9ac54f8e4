"""The identity behind the segment-parallel blend backward (csrc/blend_stream.cu, kSegRecs): the per-pixel recurrence
of the alpha-blend backward (rasterize_to_pixels_bwd.cu:192-241) can be cut at any record b if it resumes from the
forward's prefixes there -- T = T_b and B = (C_final - C_b) . v.  Pure numpy, float64 and float32."""
import numpy as np
import pytest


def _walk(alpha, c, v, v_a, t_final, lo, hi, T, B, dtype):
    out = {}
    T, B = dtype(T), dtype(B)
    for i in range(hi - 1, lo - 1, -1):
        ra = dtype(1) / (dtype(1) - alpha[i])
        T = T * ra
        fac = alpha[i] * T
        a2 = dtype(c[i] @ v)
        out[i] = T * a2 - ra * B + t_final * ra * v_a  # v_alpha of record i
        B = B + fac * a2
    return out


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-4)])
def test_backward_recurrence_can_be_cut_at_a_forward_checkpoint(dtype, tol):
    rng = np.random.default_rng(0)
    n = 300
    alpha = rng.uniform(0.004, 0.03, n).astype(dtype)
    alpha[[7, 60]] = dtype(0.9)  # large alphas: 1 / (1 - alpha) amplifies whatever the state at the cut is off by
    c = rng.uniform(0, 1, (n, 3)).astype(dtype)
    v = rng.normal(size=3).astype(dtype)
    v_a = dtype(0.3)
    T, C = dtype(1), np.zeros(3, dtype)
    t_pre, c_pre = [], []
    last = n
    for i in range(n):  # forward, with the reference's exclusive stop at T <= 1e-4
        nxt = T * (dtype(1) - alpha[i])
        if nxt <= 1e-4:
            last = i
            break
        t_pre.append(T)
        c_pre.append(C.copy())
        C = C + c[i] * (alpha[i] * T)
        T = nxt
    assert last > 40
    whole = _walk(alpha, c, v, v_a, T, 0, last, T, 0, dtype)
    for cut in (1, last // 3, last // 2, last - 1):
        back = _walk(alpha, c, v, v_a, T, cut, last, T, 0, dtype)
        front = _walk(alpha, c, v, v_a, T, 0, cut, t_pre[cut], dtype((C - c_pre[cut]) @ v), dtype)
        got = {**front, **back}
        scale = max(abs(x) for x in whole.values())
        assert max(abs(whole[i] - got[i]) for i in range(last)) <= tol * scale
