"""ctypes binding of the C ABI in include/rtp_cuda.h (realtimeparticles_b200/lib/librtp_cuda.so).

This is the only door into the compute path: there is no CPU or PyTorch fallback. If the shared library is missing
or no CUDA device is usable, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTP_CUDA_LIB") or os.path.join(_HERE, "lib", "librtp_cuda.so")  # override: A/B builds

RTP_OK, RTP_ERR_INVALID, RTP_ERR_CUDA, RTP_ERR_STATE, RTP_ERR_COMM = 0, -1, -2, -3, -4

# rtp_model
BOIDS, FLUIDS, CLOUDS = 0, 1, 2
# rtp_boundary
BOUNCING_WALL, CYCLIC_WALL = 0, 1
# rtp_step flags
STEP_PHYSICS, STEP_RENDER_AUX, STEP_CAMERA_SORT, STEP_DEBUG_FIELDS = 1, 2, 4, 8
STEP_UPDATE = STEP_PHYSICS | STEP_RENDER_AUX | STEP_CAMERA_SORT  # a full reference update()

# rtp_field
FIELDS = dict(POS=0, COL=1, VEL=2, ACC=3, PRED_POS=4, CORR_POS=5, VORT=6, TOT_CORR_POS=7, DENSITY=8,
              CONST_FACTOR=9, TEMP=10, VAPOR_DENS=11, CLOUD_DENS=12, BUOYANCY=13, CLOUD_GEN=14, PART_ID=15,
              LAPLACIAN_TEMP=16, CONST_FACTOR_TEMP=17, CORR_TEMP=18, CELL_ID=19, CAMERA_DIST=20,
              START_END_CELL=21, PERM=22, CAMERA_PERM=23, PART_DETECTOR=24)
# the reference's buffer names -> field ids
BUFFER_NAMES = {"p_pos": 0, "p_col": 1, "p_vel": 2, "p_acc": 3, "p_predPos": 4, "p_corrPos": 5, "p_vort": 6,
                "p_totCorrPos": 7, "p_density": 8, "p_constFactor": 9, "p_constFactorFld": 9, "p_temp": 10,
                "p_vaporDens": 11, "p_cloudDens": 12, "p_buoyancy": 13, "p_cloudGen": 14, "p_partID": 15,
                "p_laplacianTemp": 16, "p_constFactorTemp": 17, "p_corrTemp": 18, "p_cellID": 19,
                "p_cameraDist": 20, "c_startEndPartID": 21, "RadixSortIndices": 22, "c_partDetector": 24}
_F4 = {0, 1, 2, 3, 4, 5, 6, 7}
_U32 = {19, 20, 22, 23}


class Config(C.Structure):
    _fields_ = [("model", C.c_int32), ("device", C.c_int32), ("max_particles", C.c_uint64),
                ("nb_particles", C.c_uint64), ("box", C.c_uint32 * 3), ("grid", C.c_uint32 * 3),
                ("dim", C.c_uint32), ("max_parts_in_cell", C.c_uint32)]


class BoidsParams(C.Structure):
    """BoidsRuleKernelInputs, physics/ocl/Boids.hpp:14-20"""
    _fields_ = [("velocityScale", C.c_float), ("alignmentScale", C.c_float), ("separationScale", C.c_float),
                ("cohesionScale", C.c_float)]


class TargetParams(C.Structure):
    """TargetKernelInputs, physics/ocl/Boids.hpp:22-26"""
    _fields_ = [("targetRadiusEffect", C.c_float), ("targetSignEffect", C.c_int32)]


class FluidParams(C.Structure):
    """FluidKernelInputs, physics/ocl/Fluids.hpp:17-32"""
    _fields_ = [("restDensity", C.c_float), ("relaxCFM", C.c_float), ("timeStep", C.c_float), ("dim", C.c_uint32),
                ("isArtPressureEnabled", C.c_uint32), ("artPressureRadius", C.c_float),
                ("artPressureCoeff", C.c_float), ("artPressureExp", C.c_uint32),
                ("isVorticityConfEnabled", C.c_uint32), ("vorticityConfCoeff", C.c_float),
                ("xsphViscosityCoeff", C.c_float)]


class CloudParams(C.Structure):
    """CloudKernelInputs, physics/ocl/Clouds.hpp:15-45"""
    _fields_ = [("dim", C.c_uint32), ("timeStep", C.c_float), ("restDensity", C.c_float),
                ("groundHeatCoeff", C.c_float), ("buoyancyCoeff", C.c_float), ("gravCoeff", C.c_float),
                ("adiabaticLapseRate", C.c_float), ("phaseTransitionRate", C.c_float),
                ("latentHeatCoeff", C.c_float), ("isTempSmoothingEnabled", C.c_uint32), ("relaxCFM", C.c_float),
                ("initVaporDensityCoeff", C.c_float), ("windCoeff", C.c_float)]


assert C.sizeof(BoidsParams) == 16 and C.sizeof(TargetParams) == 8
assert C.sizeof(FluidParams) == 44 and C.sizeof(CloudParams) == 52

# every symbol include/rtp_cuda.h declares
EXPORTS = ["rtp_abi_version", "rtp_device_count", "rtp_create", "rtp_destroy", "rtp_last_error", "rtp_field_bytes",
           "rtp_upload", "rtp_download", "rtp_upload_async", "rtp_download_async", "rtp_device_ptr", "rtp_set_boids_params", "rtp_set_fluid_params",
           "rtp_set_cloud_params", "rtp_set_boundary", "rtp_set_nb_particles", "rtp_set_dimension",
           "rtp_set_displayed_quantity", "rtp_reset_ids", "rtp_init_clouds_fields", "rtp_step", "rtp_step_n",
           "rtp_sync", "rtp_get_stream", "rtp_sort_keys", "rtp_sort_keys_host", "rtp_selftest_math", "rtp_enable_profiling", "rtp_get_stage_times",
           "rtp_last_launch_count", "rtp_list_stats", "rtp_shard_set_owned", "rtp_shard_stage", "rtp_shard_buffer", "rtp_shard_list_dmax_sq", "rtp_shard_pack", "rtp_shard_unpack", "rtp_shard_clear_rows", "rtp_shard_inverse_perm", "rtp_shard_check_ghosts",
           "rtp_shard_classify", "rtp_shard_set_interior", "rtp_shard_stage_rows", "rtp_shard_exchange_stream", "rtp_shard_exchange_fork", "rtp_shard_exchange_done", "rtp_shard_exchange_join",
           "rtp_gen_box_grid", "rtp_gen_sphere_grid", "rtp_gen_rectangle_grid", "rtp_gen_circle_grid", "rtp_gen_random_box",
           "rtp_baked_constant", "rtp_register_gl", "rtp_unregister_gl", "rtp_target_create", "rtp_target_destroy", "rtp_target_update",
           "rtp_slab_group_create", "rtp_slab_group_destroy", "rtp_slab_group_last_error", "rtp_slab_group_size", "rtp_slab_group_handle",
           "rtp_slab_group_set_fluid_params", "rtp_slab_group_upload", "rtp_slab_group_step", "rtp_slab_group_sync", "rtp_slab_group_check",
           "rtp_slab_group_download"]

_lib = None


def lib():
    """Load librtp_cuda.so; fail loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "realtimeparticles_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C realtimeparticles_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32p, fp = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_float)
    L.rtp_abi_version.restype = C.c_int
    L.rtp_device_count.restype = C.c_int
    L.rtp_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.rtp_destroy.argtypes = [vp]
    L.rtp_destroy.restype = None
    L.rtp_last_error.argtypes = [vp]
    L.rtp_last_error.restype = C.c_char_p
    L.rtp_field_bytes.argtypes = [vp, C.c_int, C.POINTER(C.c_size_t)]
    L.rtp_upload.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.rtp_download.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.rtp_upload_async.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.rtp_download_async.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.rtp_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.rtp_set_boids_params.argtypes = [vp, C.POINTER(BoidsParams), C.POINTER(TargetParams), fp, C.c_int]
    L.rtp_set_fluid_params.argtypes = [vp, C.POINTER(FluidParams), C.c_int]
    L.rtp_set_cloud_params.argtypes = [vp, C.POINTER(CloudParams)]
    L.rtp_set_boundary.argtypes = [vp, C.c_int]
    L.rtp_set_nb_particles.argtypes = [vp, C.c_uint64]
    L.rtp_set_dimension.argtypes = [vp, C.c_int]
    L.rtp_set_displayed_quantity.argtypes = [vp, C.c_int, C.c_float, C.c_float]
    L.rtp_reset_ids.argtypes = [vp]
    L.rtp_init_clouds_fields.argtypes = [vp]
    L.rtp_step.argtypes = [vp, C.c_uint, fp]
    L.rtp_step_n.argtypes = [vp, C.c_uint, fp, C.c_int]
    L.rtp_sync.argtypes = [vp]
    L.rtp_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.rtp_sort_keys.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_int]
    L.rtp_sort_keys_host.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_int]
    L.rtp_selftest_math.argtypes = [vp, C.c_float, C.c_float, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.rtp_enable_profiling.argtypes = [vp, C.c_int]
    L.rtp_shard_set_owned.argtypes = [vp, C.c_uint64]
    L.rtp_shard_stage.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.rtp_shard_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.rtp_shard_pack.argtypes = [vp, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    L.rtp_shard_unpack.argtypes = [vp, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
    L.rtp_shard_clear_rows.argtypes = [vp, C.c_void_p, C.c_uint64]
    L.rtp_shard_inverse_perm.argtypes = [vp, C.c_void_p]
    L.rtp_shard_check_ghosts.argtypes = [vp, C.c_void_p, C.c_uint64, C.c_int]
    L.rtp_shard_classify.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    L.rtp_shard_set_interior.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint64]
    L.rtp_shard_stage_rows.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.rtp_shard_exchange_stream.argtypes = [vp, C.POINTER(C.c_void_p)]
    L.rtp_shard_exchange_fork.argtypes = [vp]
    L.rtp_shard_exchange_done.argtypes = [vp]
    L.rtp_shard_exchange_join.argtypes = [vp]
    L.rtp_shard_list_dmax_sq.argtypes = [vp]
    L.rtp_shard_list_dmax_sq.restype = C.c_float
    L.rtp_get_stage_times.argtypes = [vp, C.POINTER(C.c_char_p), fp, C.c_int]
    L.rtp_last_launch_count.argtypes = [vp]
    L.rtp_list_stats.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    for g in ("rtp_gen_box_grid", "rtp_gen_sphere_grid"):
        getattr(L, g).argtypes = [vp, C.POINTER(C.c_int), fp, fp]
        getattr(L, g).restype = C.c_int64
    for g in ("rtp_gen_rectangle_grid", "rtp_gen_circle_grid"):
        getattr(L, g).argtypes = [vp, C.c_int, C.POINTER(C.c_int), fp, fp]
        getattr(L, g).restype = C.c_int64
    L.rtp_gen_random_box.argtypes = [vp, C.c_int64, fp, fp, C.c_int]
    L.rtp_gen_random_box.restype = C.c_int64
    L.rtp_baked_constant.argtypes = [C.c_float]
    L.rtp_baked_constant.restype = C.c_float
    L.rtp_slab_group_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_int), C.c_uint64, u32p, u32p, C.c_uint64, C.c_uint64, C.c_int]
    L.rtp_slab_group_destroy.argtypes = [vp]
    L.rtp_slab_group_destroy.restype = None
    L.rtp_slab_group_last_error.argtypes = [vp]
    L.rtp_slab_group_last_error.restype = C.c_char_p
    L.rtp_slab_group_size.argtypes = [vp]
    L.rtp_slab_group_handle.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.rtp_slab_group_set_fluid_params.argtypes = [vp, C.POINTER(FluidParams), C.c_int]
    L.rtp_slab_group_upload.argtypes = [vp, vp, vp, C.c_uint64]
    L.rtp_slab_group_step.argtypes = [vp, C.c_int]
    L.rtp_slab_group_sync.argtypes = [vp]
    L.rtp_slab_group_check.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.rtp_slab_group_download.argtypes = [vp, vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.rtp_slab_group_download.restype = C.c_int64
    _lib = L
    return L


class RtpError(RuntimeError):
    pass


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def field_id(f):
    if isinstance(f, str):
        return BUFFER_NAMES[f] if f in BUFFER_NAMES else FIELDS[f]
    return int(f)


def _np_layout(fid, nbytes):
    if fid in _F4:
        return np.float32, (nbytes // 16, 4)
    if fid in _U32:
        return np.uint32, (nbytes // 4,)
    if fid == FIELDS["START_END_CELL"]:
        return np.uint32, (nbytes // 8, 2)
    if fid == FIELDS["PART_DETECTOR"]:
        return np.float32, (nbytes // 32, 8)
    return np.float32, (nbytes // 4,)


class Handle:
    """Owns one rtp_handle (one CUDA device + stream + all named buffers of one model)."""

    def __init__(self, model, max_particles, nb_particles, box=(10, 10, 10), grid=(30, 30, 30), dim=3,
                 max_parts_in_cell=0, device=0):
        self.L = lib()
        cfg = Config(model, device, max_particles, nb_particles, (C.c_uint32 * 3)(*box), (C.c_uint32 * 3)(*grid),
                     dim, max_parts_in_cell)
        h = C.c_void_p()
        rc = self.L.rtp_create(C.byref(cfg), C.byref(h))
        if rc != RTP_OK:
            raise RtpError("rtp_create failed (%d): %s" % (rc, (self.L.rtp_last_error(None) or b"").decode()))
        self.h = h
        self.model, self.M, self.N = model, max_particles, nb_particles
        self.ncells = grid[0] * grid[1] * grid[2]

    def close(self):
        if getattr(self, "h", None):
            self.L.rtp_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _check(self, rc, what):
        if rc != RTP_OK:
            raise RtpError("%s failed (%d): %s" % (what, rc, (self.L.rtp_last_error(self.h) or b"").decode()))

    def field_bytes(self, f):
        n = C.c_size_t()
        self._check(self.L.rtp_field_bytes(self.h, field_id(f), C.byref(n)), "rtp_field_bytes")
        return n.value

    def upload(self, f, arr, blocking=True):
        """blocking=False (rtp_upload_async): `arr` must be a page-locked, contiguous array of the field's layout that stays
        untouched until the next sync()"""
        fid = field_id(f)
        n = self.field_bytes(fid)
        dt, shape = _np_layout(fid, n)
        a = np.ascontiguousarray(np.asarray(arr, dtype=dt).reshape(shape))
        if blocking:
            self._check(self.L.rtp_upload(self.h, fid, a.ctypes.data, a.nbytes), "rtp_upload")
        else:
            self._check(self.L.rtp_upload_async(self.h, fid, a.ctypes.data, a.nbytes), "rtp_upload_async")

    def download(self, f, out=None, blocking=True):
        """blocking=False (rtp_download_async): `out` (page-locked) is complete after the next sync()"""
        fid = field_id(f)
        n = self.field_bytes(fid)
        dt, shape = _np_layout(fid, n)
        if out is None:
            out = np.empty(shape, dt)
        if blocking:
            self._check(self.L.rtp_download(self.h, fid, out.ctypes.data, out.nbytes), "rtp_download")
        else:
            self._check(self.L.rtp_download_async(self.h, fid, out.ctypes.data, out.nbytes), "rtp_download_async")
        return out

    def device_ptr(self, f):
        p = C.c_void_p()
        self._check(self.L.rtp_device_ptr(self.h, field_id(f), C.byref(p)), "rtp_device_ptr")
        return p.value

    def set_boids_params(self, rules, target=None, target_pos=None, target_active=False):
        tp = (C.c_float * 4)(*target_pos) if target_pos is not None else None
        self._check(self.L.rtp_set_boids_params(self.h, C.byref(rules), C.byref(target) if target is not None else None,
                                                tp, int(target_active)), "rtp_set_boids_params")

    def set_fluid_params(self, p, jacobi):
        self._check(self.L.rtp_set_fluid_params(self.h, C.byref(p), int(jacobi)), "rtp_set_fluid_params")

    def set_cloud_params(self, p):
        self._check(self.L.rtp_set_cloud_params(self.h, C.byref(p)), "rtp_set_cloud_params")

    def set_boundary(self, b):
        self._check(self.L.rtp_set_boundary(self.h, int(b)), "rtp_set_boundary")

    def set_nb_particles(self, n):
        self._check(self.L.rtp_set_nb_particles(self.h, int(n)), "rtp_set_nb_particles")
        self.N = int(n)

    def set_dimension(self, d):
        self._check(self.L.rtp_set_dimension(self.h, int(d)), "rtp_set_dimension")

    def set_displayed_quantity(self, f, lo, hi):
        self._check(self.L.rtp_set_displayed_quantity(self.h, field_id(f), lo, hi), "rtp_set_displayed_quantity")

    def reset_ids(self):
        self._check(self.L.rtp_reset_ids(self.h), "rtp_reset_ids")

    def init_clouds_fields(self):
        self._check(self.L.rtp_init_clouds_fields(self.h), "rtp_init_clouds_fields")

    def step(self, flags=STEP_PHYSICS, cam=(32.0, -1.2, 0.0)):
        self._check(self.L.rtp_step(self.h, flags, _f3(cam)), "rtp_step")

    def step_n(self, n, flags=STEP_PHYSICS, cam=(32.0, -1.2, 0.0)):
        self._check(self.L.rtp_step_n(self.h, flags, _f3(cam), int(n)), "rtp_step_n")

    def sync(self):
        self._check(self.L.rtp_sync(self.h), "rtp_sync")

    def stream(self):
        """cudaStream_t of the handle as an int (wrap with torch.cuda.ExternalStream to record events on it)"""
        p = C.c_void_p()
        self._check(self.L.rtp_get_stream(self.h, C.byref(p)), "rtp_get_stream")
        return p.value or 0

    def sort_keys_host(self, keys, key_bits=32):
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        out, perm = np.empty_like(keys), np.empty_like(keys)
        self._check(self.L.rtp_sort_keys_host(self.h, keys.ctypes.data, out.ctypes.data, perm.ctypes.data, keys.size,
                                              key_bits), "rtp_sort_keys_host")
        return out, perm

    def selftest_math(self, lo, hi):
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self.L.rtp_selftest_math(self.h, lo, hi, C.byref(a), C.byref(b)), "rtp_selftest_math")
        return a.value, b.value

    def shard_set_owned(self, n):
        self._check(self.L.rtp_shard_set_owned(self.h, int(n)), "rtp_shard_set_owned")

    def shard_stage(self, stage, it=0, last=False):
        self._check(self.L.rtp_shard_stage(self.h, int(stage), int(it), int(bool(last))), "rtp_shard_stage")

    def shard_buffer(self, which):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.L.rtp_shard_buffer(self.h, int(which), C.byref(p), C.byref(n)), "rtp_shard_buffer")
        return p.value, n.value

    def register_gl(self, field, vbo):
        """share an OpenGL VBO (by name) as the buffer of p_pos / p_col / c_partDetector (rtp_register_gl)"""
        self.L.rtp_register_gl.argtypes = [C.c_void_p, C.c_int, C.c_uint]
        self._check(self.L.rtp_register_gl(self.h, field_id(field), int(vbo)), "rtp_register_gl")

    def shard_pack(self, which, idx_ptr, n, out_ptr):
        self._check(self.L.rtp_shard_pack(self.h, int(which), idx_ptr, int(n), out_ptr), "rtp_shard_pack")

    def shard_unpack(self, which, idx_ptr, n, in_ptr):
        self._check(self.L.rtp_shard_unpack(self.h, int(which), idx_ptr, int(n), in_ptr), "rtp_shard_unpack")

    def shard_clear_rows(self, idx_ptr, n):
        self._check(self.L.rtp_shard_clear_rows(self.h, idx_ptr, int(n)), "rtp_shard_clear_rows")

    def shard_inverse_perm(self, out_ptr):
        self._check(self.L.rtp_shard_inverse_perm(self.h, out_ptr), "rtp_shard_inverse_perm")

    def shard_stage_rows(self, stage, it, last, rows):
        self._check(self.L.rtp_shard_stage_rows(self.h, int(stage), int(it), int(bool(last)), int(rows)), "rtp_shard_stage_rows")

    def shard_classify(self, n, layer_below, layer_from, below_ptr, above_ptr):
        self._check(self.L.rtp_shard_classify(self.h, int(n), int(layer_below), int(layer_from), below_ptr, above_ptr), "rtp_shard_classify")

    def shard_set_interior(self, cell_lo, cell_hi, max_boundary_rows):
        self._check(self.L.rtp_shard_set_interior(self.h, int(cell_lo), int(cell_hi), int(max_boundary_rows)), "rtp_shard_set_interior")

    def shard_exchange_stream(self):
        p = C.c_void_p()
        self._check(self.L.rtp_shard_exchange_stream(self.h, C.byref(p)), "rtp_shard_exchange_stream")
        return p.value

    def shard_exchange_fork(self):
        self._check(self.L.rtp_shard_exchange_fork(self.h), "rtp_shard_exchange_fork")

    def shard_exchange_done(self):
        self._check(self.L.rtp_shard_exchange_done(self.h), "rtp_shard_exchange_done")

    def shard_exchange_join(self):
        self._check(self.L.rtp_shard_exchange_join(self.h), "rtp_shard_exchange_join")

    def shard_check_ghosts(self, idx_ptr, n, next_epoch):
        self._check(self.L.rtp_shard_check_ghosts(self.h, idx_ptr, int(n), int(next_epoch)), "rtp_shard_check_ghosts")

    def shard_list_dmax_sq(self):
        return float(self.L.rtp_shard_list_dmax_sq(self.h))

    def enable_profiling(self, on):
        self._check(self.L.rtp_enable_profiling(self.h, int(on)), "rtp_enable_profiling")

    def stage_times(self):
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        n = self.L.rtp_get_stage_times(self.h, names, ms, cap)
        return [(names[i].decode(), float(ms[i])) for i in range(min(n, cap))]

    def last_launch_count(self):
        return int(self.L.rtp_last_launch_count(self.h))

    def list_stats(self):
        """Neighbour-list diagnostics after a physics-only step (rtp_list_stats)."""
        out = (C.c_ulonglong * 8)()
        self._check(self.L.rtp_list_stats(self.h, out), "rtp_list_stats")
        keys = ("particles", "margin_overflow", "cell_changed", "reserved", "sum_len", "max_len", "hit_overflow", "moved_beyond_bound")
        return dict(zip(keys, [int(v) for v in out]))


class Target:
    """Host-side boids target trajectory (rtp_target_*; replaces Physics::Target, physics/utils/Target.cpp)."""

    def __init__(self, box_size):
        L = lib()
        L.rtp_target_create.restype = C.c_void_p
        L.rtp_target_create.argtypes = [C.c_uint32]
        L.rtp_target_destroy.argtypes = [C.c_void_p]
        L.rtp_target_update.argtypes = [C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_float)]
        self.L, self.t = L, C.c_void_p(L.rtp_target_create(int(box_size)))

    def update(self, dim, velocity):
        out = (C.c_float * 3)()
        if self.L.rtp_target_update(self.t, int(dim), float(velocity), out) != 0:
            raise RtpError("rtp_target_update failed")
        return (float(out[0]), float(out[1]), float(out[2]))

    def __del__(self):
        if getattr(self, "t", None):
            self.L.rtp_target_destroy(self.t)
            self.t = None


def gen_box_grid(res, start, end):
    """Uniform lattice, x-major order (utils/Geometry.cpp:198-227); float4 rows."""
    n = res[0] * res[1] * res[2]
    out = np.empty((n, 4), np.float32)
    r = lib().rtp_gen_box_grid(out.ctypes.data, (C.c_int * 3)(*res), _f3(start), _f3(end))
    if r != n:
        raise RtpError("rtp_gen_box_grid failed: %d" % r)
    return out


def gen_sphere_grid(res, start, end):
    """Spherical lattice (utils/Geometry.cpp:243-272); float4 rows."""
    n = res[0] * res[1] * res[2]
    out = np.empty((n, 4), np.float32)
    r = lib().rtp_gen_sphere_grid(out.ctypes.data, (C.c_int * 3)(*res), _f3(start), _f3(end))
    if r != n:
        raise RtpError("rtp_gen_sphere_grid failed: %d" % r)
    return out


PLANE_XY, PLANE_XZ, PLANE_YZ = 0, 1, 2


def _gen_planar(fn, plane, res, start, end):
    n = res[0] * res[1]
    out = np.empty((n, 4), np.float32)
    r = getattr(lib(), fn)(out.ctypes.data, int(plane), (C.c_int * 2)(*res), _f3(start), _f3(end))
    if r != n:
        raise RtpError("%s failed: %d" % (fn, r))
    return out


def gen_rectangle_grid(res, start, end, plane=PLANE_YZ):
    """Planar uniform lattice of the 2D presets (utils/Geometry.cpp:75-140); float4 rows."""
    return _gen_planar("rtp_gen_rectangle_grid", plane, res, start, end)


def gen_circle_grid(res, start, end, plane=PLANE_YZ):
    """Planar polar lattice of the 2D presets (utils/Geometry.cpp:142-196); float4 rows."""
    return _gen_planar("rtp_gen_circle_grid", plane, res, start, end)


def gen_random_box(n, start, end, seed=1):
    """glibc rand() uniform fill in x,y,z call order (utils/Geometry.cpp:229-239); seed=1 == a fresh process."""
    out = np.empty((n, 4), np.float32)
    r = lib().rtp_gen_random_box(out.ctypes.data, n, _f3(start), _f3(end), seed)
    if r != n:
        raise RtpError("rtp_gen_random_box failed: %d" % r)
    return out


def baked_constant(v):
    return float(lib().rtp_baked_constant(float(v)))


class SlabGroup:
    """rtp_slab_group: the x-slab decomposition of the PBF step driven from inside the library by ONE host thread
    (csrc/slab_group.cu); slab r lives on CUDA device devices[r] (ids may repeat). Same step, same bits as
    realtimeparticles_b200.sharded.SlabDecomposition over its transports."""

    def __init__(self, devices, slab_capacity, box, grid, ghost_cap=0, migrate_cap=0, overlap=True, fluid_params=None, jacobi=3):
        self.L = lib()
        g = C.c_void_p()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = self.L.rtp_slab_group_create(C.byref(g), len(devices), devs, int(slab_capacity), (C.c_uint32 * 3)(*box), (C.c_uint32 * 3)(*grid),
                                          int(ghost_cap), int(migrate_cap), int(bool(overlap)))
        if rc != RTP_OK:
            raise RtpError("rtp_slab_group_create failed (%d): %s" % (rc, (self.L.rtp_slab_group_last_error(None) or b"").decode()))
        self.g = g
        self.n_slabs = len(devices)
        self._n = 0  # particles uploaded (sizes the download buffers)
        fp = fluid_params or FluidParams(450.0, 600.0, 0.010, 3, 1, 0.006, 0.001, 4, 1, 0.0004, 0.0001)
        self._check(self.L.rtp_slab_group_set_fluid_params(self.g, C.byref(fp), int(jacobi)), "rtp_slab_group_set_fluid_params")

    def close(self):
        if getattr(self, "g", None):
            self.L.rtp_slab_group_destroy(self.g)
            self.g = None

    def __del__(self):
        self.close()

    def _check(self, rc, what):
        if rc < 0:
            raise RtpError("%s failed (%d): %s" % (what, rc, (self.L.rtp_slab_group_last_error(self.g) or b"").decode()))
        return rc

    def upload(self, pos, vel):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        assert pos.shape == vel.shape and pos.shape[1] == 4
        self._n = len(pos)
        self._check(self.L.rtp_slab_group_upload(self.g, pos.ctypes.data, vel.ctypes.data, len(pos)), "rtp_slab_group_upload")

    def step(self, n=1):
        self._check(self.L.rtp_slab_group_step(self.g, int(n)), "rtp_slab_group_step")

    def sync(self):
        self._check(self.L.rtp_slab_group_sync(self.g), "rtp_slab_group_sync")

    def check(self):
        m = C.c_uint64(0)
        self._check(self.L.rtp_slab_group_check(self.g, C.byref(m)), "rtp_slab_group_check")
        return int(m.value)

    def download(self):
        """(pos, vel, particles per slab): all particles, slab by slab"""
        pos = np.empty((self._n, 4), np.float32)
        vel = np.empty((self._n, 4), np.float32)
        per = (C.c_uint64 * self.n_slabs)()
        n = self._check(self.L.rtp_slab_group_download(self.g, pos.ctypes.data, vel.ctypes.data, self._n, per), "rtp_slab_group_download")
        return pos[:n], vel[:n], [int(x) for x in per]
