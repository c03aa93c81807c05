"""skoots_b200 — B200-native (sm_100a) implementation of SKOOTS' skeleton-embedding
instance-assembly path behind the reference's `skoots.lib` entry points."""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
