"""chainb200 -- B200-native engine for ChainPartitioners.jl's hot path (cost oracle over sparse
column ranges + the split-point searches that consume it), behind the reference's own API.

Host side = this thin Python mirror of the Julia interface (the Julia ``ccall`` shim is in
``julia/ChainPartitionersB200.jl``); all arithmetic runs in ``libchainb200.so`` (hand-written CUDA for
sm_100a) through the C ABI declared in ``include/chainb200.h``.  There is no CPU fallback: every
entry point raises if the CUDA library is missing or no GPU is visible.
"""
from .types import *  # noqa: F401,F403
from .types import convert  # noqa: F401
from .api import *  # noqa: F401,F403
