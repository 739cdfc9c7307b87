"""B200-native mirror of the hot-path piece of the reference `utils.py`: `get_loss` (utils.py:377-387) on the fused
cross-entropy kernel.  Everything else in the reference's utils.py (logging, checkpointing, `warmup_batchnorm`, ...) is
host-side control plane that works unchanged on this model and is taken from the reference itself."""
from .engine import get_loss  # noqa: F401

__all__ = ["get_loss"]
