"""TEST INFRASTRUCTURE ONLY.  `autograd.scipy.stats` stand-in."""
from . import norm, t  # noqa: F401


class multivariate_normal(object):
    @staticmethod
    def logpdf(x, mean, cov):
        import scipy.stats as _st
        return _st.multivariate_normal.logpdf(x, mean, cov)
