"""Import shim: the package directory is `go-ctr_b200/` (not a valid Python identifier), so
`import go_ctr_b200` resolves its submodules there."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "go-ctr_b200")]

from .engine import *          # noqa: F401,F403,E402
from .model import *           # noqa: F401,F403,E402
from .item2vec import *        # noqa: F401,F403,E402
