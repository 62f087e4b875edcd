"""stand-in for PyMCubes (mesh export at the end of main_nerf.py): returns an EMPTY mesh and says so — the volumetric-rendering
path never needs it, and the reference's main must be able to run to its last line unchanged"""
import numpy as np


def marching_cubes(volume, threshold):
    print("[compat.mcubes] PyMCubes is not installed in this image: returning an empty mesh")
    return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)
