"""CPU: the C-ABI library loads and exports every symbol include/viabel_b200.h declares
(no compute calls -- there is no GPU in the build container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, 'include', 'viabel_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(vb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    lib = ctypes.CDLL(os.path.join(ROOT, 'viabel_b200', 'libviabel_b200.so'))
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), 'missing export: ' + n


def test_python_binding_covers_header():
    import viabel_b200
    assert sorted(viabel_b200._lib.EXPORTED) == _declared()
    assert viabel_b200._lib.lib.vb_version() >= 100


def test_product_never_imports_oracle():
    """The product path must not route through the CPU oracle."""
    pkg = os.path.join(ROOT, 'viabel_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('no CPU oracle', ''), f
