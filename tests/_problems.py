"""Seeded synthetic problems shared by oracle/make_golden.py and the tests.

Inputs are rebuilt from seeds on both sides (numpy legacy RandomState streams are
frozen by numpy's compatibility policy), so the goldens only need to store outputs.
"""
import numpy as np
from scipy import stats


def logistic_problem(N, d, seed):
    """X ~ N(0,1), beta* ~ N(0, 1/d), y = +1 w.p. sigmoid(x.beta*) else -1."""
    rs = np.random.RandomState(seed)
    X = rs.randn(N, d)
    beta = rs.randn(d) / np.sqrt(d)
    p = 1.0 / (1.0 + np.exp(-(X @ beta)))
    y = np.where(rs.rand(N) < p, 1.0, -1.0)
    return X, y, beta


def target_params(d, seed):
    rs = np.random.RandomState(seed)
    return rs.randn(d), np.exp(0.5 * rs.randn(d))


def hier_problem(G, p, n_per, seed):
    rs = np.random.RandomState(seed)
    N = G * n_per
    group = np.repeat(np.arange(G), n_per)
    X = rs.randn(N, p)
    m = rs.randn(p)
    beta = m + 0.5 * rs.randn(G, p)
    y = np.sum(X * beta[group], axis=1) + 0.3 * rs.randn(N)
    return dict(X=X, y=y, group=group, G=G, p=p)


def diag_problem():
    rs = np.random.RandomState(846)
    n, d = 20000, 3
    samples = rs.randn(n, d) * np.array([1.0, 2.0, 0.5]) + np.array([0.3, -1.0, 2.0])
    lw = -0.1 * np.sum(samples ** 2, axis=1) + 0.05 * rs.randn(n)
    return samples, lw


def _t_ratio(n, df_p, df_q, seed):
    rs = np.random.RandomState(seed)
    s = rs.standard_t(df_q, size=n)
    return stats.t.logpdf(s, df_p) - stats.t.logpdf(s, df_q)


PSIS_CASES = ['t5_t7_1e5', 't3_t30_2e5', 't50_t5_1e5', 'tiny_20', 'small_100', 'ties_3e4',
              'underflow_5000', 'all_equal_50', 'n2d_5000x3', 'big_1e6']


def psis_case(name):
    if name == 't5_t7_1e5':
        return _t_ratio(100000, 5, 7, 3)
    if name == 't3_t30_2e5':
        return _t_ratio(200000, 3, 30, 4)
    if name == 't50_t5_1e5':
        return _t_ratio(100000, 50, 5, 5)
    if name == 'tiny_20':
        return _t_ratio(20, 3, 30, 6)
    if name == 'small_100':
        return _t_ratio(100, 3, 9, 7)
    if name == 'ties_3e4':
        return np.tile(_t_ratio(600, 4, 9, 8), 50)
    if name == 'underflow_5000':
        rs = np.random.RandomState(9)
        return -400.0 * np.abs(rs.standard_t(2, size=5000))
    if name == 'all_equal_50':
        return np.full(50, -3.25)
    if name == 'n2d_5000x3':
        return np.stack([_t_ratio(5000, 3, 30, 10), _t_ratio(5000, 5, 7, 11),
                         _t_ratio(5000, 50, 5, 12)], axis=1)
    if name == 'big_1e6':
        return _t_ratio(1000000, 4, 9, 13)
    raise KeyError(name)
