"""esn_b200: B200-native (sm_100a) engine behind the pyESN drop-in modules in
`../libs`.  Host side is Python + torch (device memory, streams, NCCL); all
arithmetic on the hot path is hand-written CUDA reached through the C ABI of
include/esn_b200.h."""
from ._lib import EsnB200Error, LIB_PATH, load            # noqa: F401
from .engine import Reservoir                              # noqa: F401
from . import dist                                          # noqa: F401
from . import ofdm                                          # noqa: F401
