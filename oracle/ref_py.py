"""ctypes binding of oracle/_ref/*.so: the reference's OWN sources compiled for the CPU (oracle/ref/Makefile).

libref_kernels.so  : physics/ocl/kernels/*.cl executed through the OpenCL-C shim (oracle/ref/ocl_shim.hpp)
libref_geometry.so : utils/Geometry.cpp + utils/Utils.cpp, unmodified
TEST INFRASTRUCTURE ONLY: used to pin the oracle restatement (tests/test_oracle_vs_ref.py) and to generate the golden
fixtures under tests/golden/ (tests/golden/make_golden.py). Not available without a prior build in a container that
has /root/reference.
"""
import ctypes as C
import os

import numpy as np

from . import oracle_py as O

_HERE = os.path.dirname(os.path.abspath(__file__))
KERNELS = os.path.join(_HERE, "_ref", "libref_kernels.so")
GEOMETRY = os.path.join(_HERE, "_ref", "libref_geometry.so")


def available():
    return os.path.exists(KERNELS) and os.path.exists(GEOMETRY)


_k = None
_g = None


def klib():
    global _k
    if _k is None:
        L = C.CDLL(KERNELS)
        L.ref_create.argtypes = [C.POINTER(O.Config), C.POINTER(C.c_void_p)]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_field_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
        L.ref_field_ptr.restype = C.c_void_p
        L.ref_set_boids_params.argtypes = [C.c_void_p, C.POINTER(O.BoidsParams), C.POINTER(O.TargetParams),
                                           C.POINTER(C.c_float), C.c_int]
        L.ref_set_fluid_params.argtypes = [C.c_void_p, C.POINTER(O.FluidParams), C.c_int]
        L.ref_set_cloud_params.argtypes = [C.c_void_p, C.POINTER(O.CloudParams)]
        L.ref_set_boundary.argtypes = [C.c_void_p, C.c_int]
        L.ref_set_nb_particles.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_set_dimension.argtypes = [C.c_void_p, C.c_int]
        L.ref_set_displayed_quantity.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.ref_reset_ids.argtypes = [C.c_void_p]
        L.ref_init_clouds_fields.argtypes = [C.c_void_p]
        L.ref_run_stage.argtypes = [C.c_void_p, C.c_int]
        L.ref_step.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_float)]
        L.ref_constant.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_constant.restype = C.c_float
        _k = L
    return _k


def glib():
    global _g
    if _g is None:
        L = C.CDLL(GEOMETRY)
        L.ref_generate_3d_grid.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int,
                                           C.c_void_p]
        L.ref_generate_3d_grid.restype = C.c_long
        L.ref_generate_2d_grid.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int,
                                           C.c_void_p]
        L.ref_generate_2d_grid.restype = C.c_long
        L.ref_baked_constant.argtypes = [C.c_float]
        L.ref_baked_constant.restype = C.c_float
        L.ref_float_to_str.argtypes = [C.c_float, C.c_char_p, C.c_size_t]
        L.ref_srand.argtypes = [C.c_uint]
        _g = L
    return _g


class World(O.World):
    """Same interface as oracle_py.World, backed by the reference's own kernels."""

    def __init__(self, model, max_particles, nb_particles, box=(10, 10, 10), grid=(30, 30, 30), dim=3, max_parts_in_cell=0):
        self.L = _Adapter(klib())
        cfg = O.Config(model, 0, max_particles, nb_particles, (C.c_uint32 * 3)(*box), (C.c_uint32 * 3)(*grid), dim,
                       max_parts_in_cell)
        h = C.c_void_p()
        rc = klib().ref_create(C.byref(cfg), C.byref(h))
        assert rc == 0
        self.h = h
        self.model, self.M, self.N = model, max_particles, nb_particles
        self.ncells = grid[0] * grid[1] * grid[2]
        klib().ref_reset_ids(self.h)


class _Adapter:
    """maps the orc_* names used by oracle_py.World onto the ref_* entry points"""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        if name.startswith("orc_"):
            return getattr(self._lib, "ref_" + name[4:])
        raise AttributeError(name)


def generate_3d_grid(shape, res, start, end, random=False, seed=None):
    """Geometry::Generate3DGrid (utils/Geometry.cpp:40-75) of the reference itself; shape 0 = box, 1 = sphere."""
    n = res[0] * res[1] * res[2]
    out = np.empty((n, 4), np.float32)
    if seed is not None:
        glib().ref_srand(seed)
    r = glib().ref_generate_3d_grid(shape, (C.c_int * 3)(*res), (C.c_float * 3)(*start), (C.c_float * 3)(*end), int(random),
                                    out.ctypes.data)
    assert r == n
    return out


def generate_2d_grid(shape, plane, res, start, end, random=False, seed=None):
    """Geometry::Generate2DGrid (utils/Geometry.cpp:8-38) of the reference itself; shape 0 = rectangle, 1 = circle;
    plane 0 = XY, 1 = XZ, 2 = YZ."""
    n = res[0] * res[1]
    out = np.empty((n, 4), np.float32)
    if seed is not None:
        glib().ref_srand(seed)
    r = glib().ref_generate_2d_grid(shape, plane, (C.c_int * 2)(*res), (C.c_float * 3)(*start), (C.c_float * 3)(*end),
                                    int(random), out.ctypes.data)
    assert r == n
    return out


def baked_constant(v):
    return float(glib().ref_baked_constant(float(v)))


def float_to_str(v):
    buf = C.create_string_buffer(64)
    glib().ref_float_to_str(float(v), buf, 64)
    return buf.value.decode()


def target_trajectory(box_size, dim, velocity, steps):
    """positions of the reference's own Physics::Target (physics/utils/Target.cpp + PerlinNoise.cpp) over `steps` updates"""
    L = glib()
    L.ref_target_create.restype = C.c_void_p
    L.ref_target_create.argtypes = [C.c_uint]
    L.ref_target_destroy.argtypes = [C.c_void_p]
    L.ref_target_update.argtypes = [C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_float)]
    t = C.c_void_p(L.ref_target_create(int(box_size)))
    out = np.empty((steps, 3), np.float32)
    buf = (C.c_float * 3)()
    for k in range(steps):
        L.ref_target_update(t, int(dim), float(velocity), buf)
        out[k] = (buf[0], buf[1], buf[2])
    L.ref_target_destroy(t)
    return out
