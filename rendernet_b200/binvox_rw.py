"""NumPy-2-safe mirror of the reader half of the reference's tools/binvox_rw.py (:45-93)."""
from __future__ import annotations

import numpy as np


class Voxels(object):
    """:10-43."""

    def __init__(self, data, dims, translate, scale, axis_order):
        self.data = data
        self.dims = dims
        self.translate = translate
        self.scale = scale
        assert axis_order in ('xzy', 'xyz')
        self.axis_order = axis_order

    def clone(self):
        return Voxels(self.data.copy(), self.dims[:], self.translate[:], self.scale, self.axis_order)


def read_header(fp):
    """:45-56."""
    line = fp.readline().strip()
    if not line.startswith(b'#binvox'):
        raise IOError('Not a binvox file')
    dims = list(map(int, fp.readline().strip().split(b' ')[1:]))
    translate = list(map(float, fp.readline().strip().split(b' ')[1:]))
    scale = list(map(float, fp.readline().strip().split(b' ')[1:]))[0]
    fp.readline()
    return dims, translate, scale


def read_as_3d_array(fp, fix_coords=True):
    """:58-93: RLE (value,count) byte pairs -> bool[dims]; xzy -> xyz transpose when fix_coords."""
    dims, translate, scale = read_header(fp)
    raw_data = np.frombuffer(fp.read(), dtype=np.uint8)
    values, counts = raw_data[::2], raw_data[1::2]
    data = np.repeat(values, counts).astype(bool)
    data = data.reshape(dims)
    if fix_coords:
        data = np.transpose(data, (0, 2, 1))
        axis_order = 'xyz'
    else:
        axis_order = 'xzy'
    return Voxels(data, dims, translate, scale, axis_order)
