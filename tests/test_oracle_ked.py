"""Known-answer tests that pin the oracle's restatement of gstat's kriging with external drift
(no gstat here, so analytic identities are the pin).  CPU only."""
import numpy as np
import pytest

from oracle import twx_oracle as o


def _case(seed, n=60):
    r = np.random.default_rng(seed)
    lon, lat = r.uniform(-101, -97, n), r.uniform(38, 43, n)
    elev, lst = r.uniform(200, 2500, n), r.normal(10, 6, n)
    X = np.column_stack([lon, lat, elev, lst])
    y = 12 - 0.005 * elev - 0.5 * (lat - 40) + 0.2 * lst + r.normal(0, 0.7, n)
    pt = np.array([-99.1, 40.4, 950.0, 11.0])
    return lon, lat, X, y, pt


def test_distance_known_answers():
    assert abs(o.gcdist_sp(0, 0, 1, 0) - 111.3195) < 1e-4          # 1 deg of longitude on the equator
    assert abs(o.gcdist_sp(0, 0, 0, 1) - 110.5731) < 1e-4          # 0 -> 1 deg along a meridian
    assert o.gcdist_sp(-100.0, 40.0, -100.0, 40.0) == 0.0
    assert abs(o.grt_circle_dist(0.0, 0.0, 1.0, 0.0) - 111.1951) < 1e-4
    # symmetric
    a = o.gcdist_sp(-100.3, 40.2, -99.1, 41.7)
    assert a == o.gcdist_sp(-99.1, 41.7, -100.3, 40.2)


def test_pure_nugget_is_ols():
    lon, lat, X, y, pt = _case(1)
    nug, psill = 0.3, 1.2
    mean, var = o.ked_gstat(lon, lat, X, y, pt[0], pt[1], pt, nug, psill, 0.0)
    A = np.column_stack([np.ones(len(y)), X])
    x0 = np.concatenate([[1.0], pt])
    b, *_ = np.linalg.lstsq(A, y, rcond=None)
    s = nug + psill
    # centre for a stable closed form of x0'(X'X)^-1 x0
    Ac = np.column_stack([np.ones(len(y)), X - pt])
    q = np.linalg.solve(Ac.T @ Ac, np.eye(5)[0])[0]
    assert abs(mean - x0 @ b) < 1e-8
    assert abs(var - s * (1 + q)) < 1e-10


def test_exact_at_data_location():
    lon, lat, X, y, _ = _case(2)
    j = 7
    mean, var = o.ked_gstat(lon, lat, X, y, lon[j], lat[j], X[j], 0.2, 1.5, 80.0)
    assert abs(mean - y[j]) < 1e-9
    assert abs(var) < 1e-9


def test_exact_linear_field():
    lon, lat, X, _, pt = _case(3)
    b = np.array([0.3, -0.6, -0.006, 0.25])
    y = 4.0 + X @ b
    mean, var = o.ked_gstat(lon, lat, X, y, pt[0], pt[1], pt, 0.1, 2.0, 120.0)
    assert abs(mean - (4.0 + pt @ b)) < 1e-8
    assert var > 0


def test_matches_augmented_system_in_extended_precision():
    """(n+5)x(n+5) universal-kriging system, uncentred drift, solved in longdouble == GLS form."""
    lon, lat, X, y, pt = _case(4, n=45)
    nug, psill, rng = 0.15, 1.1, 60.0
    mean, var = o.ked_gstat(lon, lat, X, y, pt[0], pt[1], pt, nug, psill, rng)
    n = len(y)
    H = o.gcdist_sp(lon[:, None], lat[:, None], lon[None, :], lat[None, :])
    V = o.covariance_gstat(H, nug, psill, rng)
    c0 = o.covariance_gstat(o.gcdist_sp(pt[0], pt[1], lon, lat), nug, psill, rng)
    ld = np.longdouble
    # scale drift columns (pure reparametrisation) so the longdouble elimination stays accurate
    sc = np.array([1.0, 100.0, 40.0, 1000.0, 10.0])
    F = np.column_stack([np.ones(n), X]) / sc
    f0 = np.concatenate([[1.0], pt]) / sc
    K = np.zeros((n + 5, n + 5), dtype=ld)
    K[:n, :n] = V
    K[:n, n:] = F
    K[n:, :n] = F.T
    rhs = np.concatenate([c0, f0]).astype(ld)
    # Gaussian elimination with partial pivoting in longdouble
    A = np.column_stack([K, rhs])
    m = n + 5
    for k in range(m):
        p = k + int(np.argmax(np.abs(A[k:, k])))
        A[[k, p]] = A[[p, k]]
        A[k + 1:] -= np.outer(A[k + 1:, k] / A[k, k], A[k])
    sol = np.zeros(m, dtype=ld)
    for k in range(m - 1, -1, -1):
        sol[k] = (A[k, m] - A[k, k + 1:m] @ sol[k + 1:]) / A[k, k]
    lam, mu = sol[:n], sol[n:]
    mean_uk = float(lam @ y.astype(ld))
    var_uk = float((nug + psill) - lam @ c0.astype(ld) - mu @ f0.astype(ld))
    assert abs(mean - mean_uk) < 1e-7
    assert abs(var - var_uk) < 1e-7 * max(1.0, abs(var_uk))
    assert abs(float(lam.sum()) - 1.0) < 1e-9


def test_singular_duplicate_locations():
    lon, lat, X, y, pt = _case(5)
    lon[3], lat[3] = lon[9], lat[9]
    with pytest.raises(o.OracleError) as e:
        o.ked_gstat(lon, lat, X, y, pt[0], pt[1], pt, 0.2, 1.0, 50.0)
    assert e.value.status == o.ST_SINGULAR
