"""viabel_b200 -- B200-native engine for viabel's data-parallel hot path.

Same public names as `viabel` (viabel/__init__.py:1-6) for the in-scope path.  Importing the
package loads libviabel_b200.so and fails loudly if it has not been built.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from ._psis import *  # noqa: F401,F403
from .approximations import *  # noqa: F401,F403
from .convenience import *  # noqa: F401,F403
from .diagnostics import *  # noqa: F401,F403
from .flows import *  # noqa: F401,F403
from .models import *  # noqa: F401,F403
from .objectives import *  # noqa: F401,F403
from .optimization import *  # noqa: F401,F403

__version__ = '0.1.0'
