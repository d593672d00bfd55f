"""ocelot_b200 -- B200-native 3D space-charge kick behind Ocelot's PhysProc API."""
from .physproc import PhysProc
from .particles import ParticleArray, DeviceParticleArray
from .sc import SpaceCharge
from .sc import install as install_space_charge, uninstall as uninstall_space_charge
from .lsc import LSC
from .lsc import install as install_lsc, uninstall as uninstall_lsc
from .beam import apply_map, get_envelope, Moments, EnvelopeRecorder
from .track import track, replay_track
from .apertures import RectAperture, EllipticalAperture
from .io import save_particle_array2npz, load_particle_array_from_npz

__all__ = ["PhysProc", "ParticleArray", "DeviceParticleArray", "SpaceCharge", "LSC", "install", "uninstall", "install_space_charge", "install_lsc",
           "apply_map", "get_envelope", "Moments", "EnvelopeRecorder", "track", "replay_track",
           "RectAperture", "EllipticalAperture", "save_particle_array2npz", "load_particle_array_from_npz"]


def install():
    """Replace Ocelot's ``SpaceCharge`` and ``LSC`` (``ocelot.cpbd.sc`` and the ``ocelot`` re-exports) with the
    B200 classes, so unmodified Ocelot scripts pick them up.  Returns the two classes."""
    return install_space_charge(), install_lsc()


def uninstall():
    """Undo ``install()``: every rebound name gets the reference class back."""
    uninstall_space_charge()
    uninstall_lsc()
