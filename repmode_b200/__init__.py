"""repmode_b200: B200-native (sm_100a) implementation of RepMode's MoDE-conv hot path behind the reference's
`fnet.nn_modules.RepMode` plugin API.  See DESIGN.md."""
from . import lib  # noqa: F401

__all__ = ["lib"]
