"""TEST INFRASTRUCTURE ONLY.  `autograd.extend` stand-in (models.py:1 imports it)."""


def primitive(f):
    return f


def defvjp(*a, **k):
    return None
