"""Host side of the quantised support sampling (``sampling_quantized``,
source/poco_data_loader.py:59-134): the random rotations.  The reference rotates the cloud by a random angle in
[-180, 180] degrees about x, then y, then z before every voxel-grid round (poco_data_loader.py:90-103); the angles are
drawn here from a numpy generator (a seed reproduces a sampling) and shipped to the device as 3x3 matrices, everything
else runs in ``csrc/sampling.cu``."""
import math

import numpy as np

ROUNDS = 6  # voxel-halving rounds the device sampler is given rotations for (two or three are used in practice)


def random_rotation(gen: np.random.Generator) -> np.ndarray:
    mats = []
    for axis in range(3):
        deg = gen.uniform(-180.0, 180.0)
        c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
        m = np.eye(3)
        a, b = [(1, 2), (0, 2), (0, 1)][axis]
        m[a, a], m[a, b], m[b, a], m[b, b] = c, s, -s, c
        mats.append(m)
    return mats[2] @ mats[1] @ mats[0]  # rot_z(rot_y(rot_x(.)))


def random_rotations(gen: np.random.Generator, count: int) -> np.ndarray:
    """``[count, 9]`` float32 row-major rotation matrices (vectorised ``random_rotation``: same angles in the same order)"""
    ang = np.radians(gen.uniform(-180.0, 180.0, (count, 3)))
    c, s = np.cos(ang), np.sin(ang)
    one, zero = np.ones(count), np.zeros(count)

    def mat(rows):
        return np.stack([np.stack(r, axis=-1) for r in rows], axis=-2)

    rx = mat([[one, zero, zero], [zero, c[:, 0], s[:, 0]], [zero, -s[:, 0], c[:, 0]]])
    ry = mat([[c[:, 1], zero, s[:, 1]], [zero, one, zero], [-s[:, 1], zero, c[:, 1]]])
    rz = mat([[c[:, 2], s[:, 2], zero], [-s[:, 2], c[:, 2], zero], [zero, zero, one]])
    return (rz @ ry @ rx).reshape(count, 9).astype(np.float32)
