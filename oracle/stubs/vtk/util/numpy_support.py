import numpy as np


def numpy_to_vtk(a, *args, **kwargs):
    return np.asarray(a)


def vtk_to_numpy(a):
    return np.asarray(a)
