"""TEST INFRASTRUCTURE ONLY.  Empty stand-in: viabel/_utils.py:8 imports pystan at
module import time; nothing on the in-scope path calls it."""


def StanModel(*a, **k):
    raise RuntimeError('pystan is not available in this image')
