"""qcdenoise_b200 -- B200-native noisy-circuit sampling engine behind qcdenoise's samplers / stabilizers /
witnesses / simulate API.  The simulation path lives in libqcdsim.so (hand-written sm_100a CUDA behind a C ABI);
importing this package fails if that library has not been built."""
from . import _lib

_lib.load()

from .ir import Circuit, encode_programs  # noqa: E402
from .engine import Engine, debug_schedule  # noqa: E402
from .graph_circuit import *  # noqa: E402,F401,F403
from .graph_state import *  # noqa: E402,F401,F403
from .samplers import *  # noqa: E402,F401,F403
from .stabilizers import *  # noqa: E402,F401,F403
from .witnesses import *  # noqa: E402,F401,F403
from .simulate import *  # noqa: E402,F401,F403
from .dataset import *  # noqa: E402,F401,F403
from . import batch, channels, distributed  # noqa: E402,F401
