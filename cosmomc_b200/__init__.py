"""cosmomc_b200 — B200-native (sm_100a) theory + likelihood hot path behind CosmoMC's calculator / likelihood
plug-in surface.  See DESIGN.md; the C ABI is include/cosmob200.h, loaded here through ctypes (lib.py)."""
from . import lib  # noqa: F401

__all__ = ["lib"]


def __getattr__(name):
    # `cosmomc_b200.LikeCalculator(ini)`: SURVEY 8b's Python entry (lazy: keeps `import cosmomc_b200` free of side effects)
    if name == "LikeCalculator":
        from .likecalc import LikeCalculator
        return LikeCalculator
    raise AttributeError(name)
