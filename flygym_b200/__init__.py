"""flygym_b200 — B200-native batched stepper for NeuroMechFly.

Drop-in for the ``step()/reset()/get_*()/set_*()`` surface of the reference's
``flygym.Simulation`` / ``flygym.warp.GPUSimulation``.  See DESIGN.md.
"""
from .anatomy import ActuatorType  # noqa: F401
from .model import NMFModel  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # heavy (torch / ctypes) pieces are imported lazily
    if name == "B200Simulation":
        from .simulation import B200Simulation
        return B200Simulation
    if name == "NMFVectorEnv":
        from .vecenv import NMFVectorEnv
        return NMFVectorEnv
    if name in ("Retina",):
        from .retina import Retina
        return Retina
    raise AttributeError(name)
