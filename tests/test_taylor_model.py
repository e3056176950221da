"""The bound behind the Taylor-model optimiser (bito_b200/csrc/gp_types.h, OptPass), checked in numpy at
extended precision: for S(x) = sum_p w_p log(1 + rho_p x), z_p = rho_p / (1 + rho_p c), M_j = sum_p w_p z_p^j,

    | S(x) - S(c) - sum_{j < J} (-1)^{j+1} M_j (x - c)^j / j |  <=  2 M_J |x - c|^J / J     (J even)

whenever max_p |z_p (x - c)| <= 1/2, and max_p |z_p|^J <= M_J / min_p w_p. This is host-side arithmetic on
the formulas only (no engine, no oracle): the CUDA kernels are checked against the oracle in the GPU tests."""
import numpy as np

J = 12


def _model(rho, w, c):
    z = rho / (1.0 + rho * c)
    return np.array([np.sum(w * z ** j) for j in range(1, J + 1)], dtype=np.longdouble), z


def test_truncation_bound_holds():
    rng = np.random.default_rng(5)
    for trial in range(200):
        n = int(rng.integers(10, 4000))
        # JC69 ratios live in [-1, 3]; a few patterns close to the singular end rho = -1
        rho = rng.uniform(-0.9, 3.0, n).astype(np.longdouble)
        k = max(1, n // 50)
        rho[:k] = -1.0 + rng.uniform(1e-6, 5e-2, k)
        w = rng.integers(1, 6, n).astype(np.longdouble)
        c = np.longdouble(rng.uniform(0.3, 0.999))
        M, z = _model(rho, w, c)
        zmax = np.max(np.abs(z))
        assert zmax ** J <= M[J - 1] / np.min(w) * (1 + 1e-12)
        for frac in (0.5, 0.3, 0.1, 0.01):
            d = np.longdouble(frac) / zmax * (1 if trial % 2 else -1)
            x = c + d
            if not (0.0 < x <= 1.0):
                continue
            exact = np.sum(w * (np.log1p(rho * x) - np.log1p(rho * c)))
            series = sum((-1) ** (j + 1) * M[j - 1] * d ** j / j for j in range(1, J))
            bound = 2 * M[J - 1] * abs(d) ** J / J
            noise = 64 * np.finfo(np.longdouble).eps * np.sum(w * np.abs(np.log1p(rho * x)))
            assert abs(exact - series) <= bound + noise, (trial, frac, float(abs(exact - series)), float(bound))


def test_newton_identities_match_direct_power_sums():
    """k_opt_eval_model forms the power sums of four values from their elementary symmetric polynomials."""
    rng = np.random.default_rng(6)
    for _ in range(500):
        a, b, c, d = (rng.uniform(-3, 3, 4) if rng.random() < 0.5 else 0.83 + 1e-4 * rng.standard_normal(4))
        s_ab, s_cd, p_ab, p_cd = a + b, c + d, a * b, c * d
        e1, e2, e3, e4 = s_ab + s_cd, s_ab * s_cd + (p_ab + p_cd), p_ab * s_cd + p_cd * s_ab, p_ab * p_cd
        p = [e1, a * a + b * b + c * c + d * d]
        p.append(e1 * p[1] - e2 * p[0] + 3 * e3)
        p.append(e1 * p[2] - e2 * p[1] + e3 * p[0] - 4 * e4)
        for j in range(4, J):
            p.append(e1 * p[j - 1] - e2 * p[j - 2] + e3 * p[j - 3] - e4 * p[j - 4])
        z = np.array([a, b, c, d], dtype=np.longdouble)
        for j in range(1, J + 1):
            want = np.sum(z ** j)
            scale = np.sum(np.abs(z) ** j)
            assert abs(p[j - 1] - want) <= 1e-11 * scale
