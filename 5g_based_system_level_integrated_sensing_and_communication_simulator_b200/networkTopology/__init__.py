"""``+networkTopology`` package mirror (only the part next to the hot path: the LoS / blockage geometry)."""
from . import blockages  # noqa: F401
