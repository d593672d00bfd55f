"""ocelot_b200 -- B200-native 3D space-charge kick behind Ocelot's PhysProc API."""
from .physproc import PhysProc
from .particles import ParticleArray, DeviceParticleArray
from .sc import SpaceCharge, install
from .lsc import LSC
from .beam import apply_map, get_envelope, Moments
from .track import track, replay_track
from .apertures import RectAperture, EllipticalAperture
from .io import save_particle_array2npz, load_particle_array_from_npz

__all__ = ["PhysProc", "ParticleArray", "DeviceParticleArray", "SpaceCharge", "LSC", "install",
           "apply_map", "get_envelope", "Moments", "track", "replay_track",
           "RectAperture", "EllipticalAperture", "save_particle_array2npz", "load_particle_array_from_npz"]
