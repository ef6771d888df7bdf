"""ctypes binding of the CPU oracle (oracle/librtp_oracle.so). TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only;
nothing under realtimeparticles_b200/ imports it. See oracle/rtp_oracle.h for the parity-pin statement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librtp_oracle.so")

# rtp_field ids (include/rtp_cuda.h)
F = dict(POS=0, COL=1, VEL=2, ACC=3, PRED_POS=4, CORR_POS=5, VORT=6, TOT_CORR_POS=7, DENSITY=8, CONST_FACTOR=9,
         TEMP=10, VAPOR_DENS=11, CLOUD_DENS=12, BUOYANCY=13, CLOUD_GEN=14, PART_ID=15, LAPLACIAN_TEMP=16,
         CONST_FACTOR_TEMP=17, CORR_TEMP=18, CELL_ID=19, CAMERA_DIST=20, START_END_CELL=21, PERM=22,
         CAMERA_PERM=23, PART_DETECTOR=24)
_F4 = {0, 1, 2, 3, 4, 5, 6, 7}
_U32 = {19, 20, 22, 23}

# orc_stage ids (oracle/rtp_oracle.h)
STAGES = ["FILL_CELL_IDS", "SORT_BY_CELL", "BUILD_CELL_TABLE", "BD_RULES", "BD_TARGET", "BD_UPDATE_VEL",
          "BD_UPDATE_POS", "PREDICT_POS", "APPLY_BOUNDARY", "DENSITY", "CONSTRAINT_FACTOR", "CONSTRAINT_CORRECTION",
          "CORRECT_POS", "UPDATE_VEL", "VORTICITY", "VORTICITY_CONFINEMENT", "XSPH", "UPDATE_POS", "CLD_THERMO",
          "CLD_LAPLACIAN_TEMP", "CLD_CONSTRAINT_FACTOR_TEMP", "CLD_CONSTRAINT_CORRECTION_TEMP", "CLD_CORRECT_TEMP",
          "RENDER_AUX", "CAMERA_SORT"]
S = {n: i for i, n in enumerate(STAGES)}

BOIDS, FLUIDS, CLOUDS = 0, 1, 2
STEP_PHYSICS, STEP_RENDER_AUX, STEP_CAMERA_SORT, STEP_DEBUG_FIELDS = 1, 2, 4, 8


class Config(C.Structure):
    _fields_ = [("model", C.c_int32), ("device", C.c_int32), ("max_particles", C.c_uint64),
                ("nb_particles", C.c_uint64), ("box", C.c_uint32 * 3), ("grid", C.c_uint32 * 3),
                ("dim", C.c_uint32), ("max_parts_in_cell", C.c_uint32)]


class BoidsParams(C.Structure):
    _fields_ = [("velocityScale", C.c_float), ("alignmentScale", C.c_float), ("separationScale", C.c_float),
                ("cohesionScale", C.c_float)]


class TargetParams(C.Structure):
    _fields_ = [("targetRadiusEffect", C.c_float), ("targetSignEffect", C.c_int32)]


class FluidParams(C.Structure):
    _fields_ = [("restDensity", C.c_float), ("relaxCFM", C.c_float), ("timeStep", C.c_float), ("dim", C.c_uint32),
                ("isArtPressureEnabled", C.c_uint32), ("artPressureRadius", C.c_float),
                ("artPressureCoeff", C.c_float), ("artPressureExp", C.c_uint32),
                ("isVorticityConfEnabled", C.c_uint32), ("vorticityConfCoeff", C.c_float),
                ("xsphViscosityCoeff", C.c_float)]


class CloudParams(C.Structure):
    _fields_ = [("dim", C.c_uint32), ("timeStep", C.c_float), ("restDensity", C.c_float),
                ("groundHeatCoeff", C.c_float), ("buoyancyCoeff", C.c_float), ("gravCoeff", C.c_float),
                ("adiabaticLapseRate", C.c_float), ("phaseTransitionRate", C.c_float),
                ("latentHeatCoeff", C.c_float), ("isTempSmoothingEnabled", C.c_uint32), ("relaxCFM", C.c_float),
                ("initVaporDensityCoeff", C.c_float), ("windCoeff", C.c_float)]


def default_fluid_params(dim=3):
    return FluidParams(450.0, 600.0, 0.010, dim, 1, 0.006, 0.001, 4, 1, 0.0004, 0.0001)


def default_cloud_params(dim=3):
    return CloudParams(dim, 0.01, 450.0, 10.0, 0.10, 0.0005, 5.0, 0.3485, 0.07, 1, 600.0, 0.75, 1.0)


def default_boids_params():
    return BoidsParams(0.5, 1.6, 1.6, 1.45)


def build(force=False):
    """Compile oracle/librtp_oracle.so with the committed Makefile (building the checker is not using it)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "rtp_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_field_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
        L.orc_field_ptr.restype = C.c_void_p
        L.orc_set_boids_params.argtypes = [C.c_void_p, C.POINTER(BoidsParams), C.POINTER(TargetParams),
                                           C.POINTER(C.c_float), C.c_int]
        L.orc_set_fluid_params.argtypes = [C.c_void_p, C.POINTER(FluidParams), C.c_int]
        L.orc_set_cloud_params.argtypes = [C.c_void_p, C.POINTER(CloudParams)]
        L.orc_set_boundary.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_nb_particles.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_set_dimension.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_displayed_quantity.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.orc_set_camera.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.orc_reset_ids.argtypes = [C.c_void_p]
        L.orc_init_clouds_fields.argtypes = [C.c_void_p]
        L.orc_run_stage.argtypes = [C.c_void_p, C.c_int]
        L.orc_step.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_float)]
        L.orc_constant.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_constant.restype = C.c_float
        L.orc_baked_constant.argtypes = [C.c_float]
        L.orc_baked_constant.restype = C.c_float
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_sort_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        for g in ("orc_gen_box_grid", "orc_gen_sphere_grid"):
            getattr(L, g).argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float)]
            getattr(L, g).restype = C.c_int64
        L.orc_gen_random_box.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]
        L.orc_gen_random_box.restype = C.c_int64
        _lib = L
    return _lib


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class World:
    """One oracle instance == one reference model object (state + parameters)."""

    def __init__(self, model, max_particles, nb_particles, box=(10, 10, 10), grid=(30, 30, 30), dim=3,
                 max_parts_in_cell=0):
        self.L = lib()
        cfg = Config(model, 0, max_particles, nb_particles, (C.c_uint32 * 3)(*box), (C.c_uint32 * 3)(*grid), dim,
                     max_parts_in_cell)
        h = C.c_void_p()
        rc = self.L.orc_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise ValueError("orc_create failed: %d" % rc)
        self.h = h
        self.model, self.M, self.N = model, max_particles, nb_particles
        self.ncells = grid[0] * grid[1] * grid[2]
        self.L.orc_reset_ids(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def field(self, fid):
        """numpy view (no copy) of a named buffer."""
        if isinstance(fid, str):
            fid = F[fid]
        n = C.c_size_t()
        p = self.L.orc_field_ptr(self.h, fid, C.byref(n))
        if not p:
            raise KeyError(fid)
        if fid in _F4:
            dt, shape = np.float32, (n.value // 16, 4)
        elif fid in _U32:
            dt, shape = np.uint32, (n.value // 4,)
        elif fid == F["START_END_CELL"]:
            dt, shape = np.uint32, (n.value // 8, 2)
        elif fid == F["PART_DETECTOR"]:
            dt, shape = np.float32, (n.value // 32, 8)
        else:
            dt, shape = np.float32, (n.value // 4,)
        buf = (C.c_char * n.value).from_address(p)
        return np.frombuffer(buf, dtype=dt).reshape(shape)

    def upload(self, fid, arr):
        v = self.field(fid)
        v[...] = np.asarray(arr, dtype=v.dtype).reshape(v.shape)

    def download(self, fid):
        return self.field(fid).copy()

    def set_fluid_params(self, p, jacobi):
        self.L.orc_set_fluid_params(self.h, C.byref(p), jacobi)

    def set_cloud_params(self, p):
        self.L.orc_set_cloud_params(self.h, C.byref(p))

    def set_boids_params(self, rules, target=None, target_pos=None, target_active=False):
        tp = (C.c_float * 4)(*target_pos) if target_pos is not None else None
        self.L.orc_set_boids_params(self.h, C.byref(rules), C.byref(target) if target is not None else None, tp,
                                    int(target_active))

    def set_boundary(self, b):
        self.L.orc_set_boundary(self.h, b)

    def set_nb_particles(self, n):
        self.L.orc_set_nb_particles(self.h, n)
        self.N = n

    def set_dimension(self, d):
        self.L.orc_set_dimension(self.h, d)

    def set_displayed_quantity(self, fid, lo, hi):
        self.L.orc_set_displayed_quantity(self.h, fid, lo, hi)

    def reset_ids(self):
        self.L.orc_reset_ids(self.h)

    def init_clouds_fields(self):
        rc = self.L.orc_init_clouds_fields(self.h)
        assert rc == 0

    def run_stage(self, name):
        rc = self.L.orc_run_stage(self.h, S[name])
        assert rc == 0, name

    def step(self, flags=STEP_PHYSICS, cam=(32.0, -1.2, 0.0)):
        rc = self.L.orc_step(self.h, flags, _f3(cam))
        assert rc == 0

    def constant(self, name):
        return float(self.L.orc_constant(self.h, name.encode()))


def sort_keys(keys):
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    out = np.empty_like(keys)
    perm = np.empty_like(keys)
    lib().orc_sort_keys(keys.ctypes.data, out.ctypes.data, perm.ctypes.data, keys.size)
    return out, perm


def gen_box_grid(res, start, end):
    n = res[0] * res[1] * res[2]
    out = np.empty((n, 4), np.float32)
    lib().orc_gen_box_grid(out.ctypes.data, (C.c_int * 3)(*res), _f3(start), _f3(end))
    return out


def gen_sphere_grid(res, start, end):
    n = res[0] * res[1] * res[2]
    out = np.empty((n, 4), np.float32)
    lib().orc_gen_sphere_grid(out.ctypes.data, (C.c_int * 3)(*res), _f3(start), _f3(end))
    return out


def gen_random_box(n, start, end, seed=1):
    out = np.empty((n, 4), np.float32)
    lib().orc_gen_random_box(out.ctypes.data, n, _f3(start), _f3(end), seed)
    return out


def max_threads():
    return int(lib().orc_max_threads())
