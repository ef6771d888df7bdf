"""ctypes wrapper of oracle/_ref/libref_ocl.so: the reference's unmodified .cl kernels on the box's OpenCL GPU device
(oracle/ocl/ref_ocl.c). TEST INFRASTRUCTURE ONLY (tests/, bench.py's cpu_baseline leg)."""
import ctypes as C
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_ocl.so")
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.rocl_create.restype = C.c_void_p
        L.rocl_create.argtypes = [C.c_uint, C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.c_int, C.c_char_p, C.c_size_t]
        L.rocl_destroy.argtypes = [C.c_void_p]
        L.rocl_upload.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        L.rocl_download.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        L.rocl_step.argtypes = [C.c_void_p, C.c_void_p]
        L.rocl_reset_ids.argtypes = [C.c_void_p]
        L.rocl_set_fluid_params.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.rocl_last_error.restype = C.c_char_p
        L.rocl_last_error.argtypes = [C.c_void_p]
        L.rocl_device_name.restype = C.c_char_p
        L.rocl_device_name.argtypes = [C.c_void_p]
        L.rocl_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_uint), C.c_int, C.c_int]
        _lib = L
    return _lib


class OclFluids:
    """Fluids model of the reference on the OpenCL GPU device: physics part of Fluids::update()."""

    SHAPES = {"p_pos": 4, "p_vel": 4, "p_predPos": 4, "p_corrPos": 4, "p_vort": 4, "p_density": 1, "p_constFactor": 1}

    def __init__(self, M, N, box=(10, 10, 10), grid=(30, 30, 30), jacobi=3):
        err = C.create_string_buffer(2048)
        self.L, self.M, self.N, self.cells = lib(), M, N, grid[0] * grid[1] * grid[2]
        self.h = self.L.rocl_create(M, N, (C.c_uint * 3)(*box), (C.c_uint * 3)(*grid), jacobi, err, 2048)
        if not self.h:
            raise RuntimeError("reference OpenCL runner: " + err.value.decode(errors="replace"))
        self.h = C.c_void_p(self.h)
        self.device = self.L.rocl_device_name(self.h).decode()

    def upload(self, name, arr):
        a = np.ascontiguousarray(arr)
        if self.L.rocl_upload(self.h, name.encode(), a.ctypes.data, a.nbytes) != 0:
            raise RuntimeError("rocl_upload " + name)

    def download(self, name):
        if name == "p_cellID":
            out = np.empty(self.M, np.uint32)
        elif name == "c_startEndPartID":
            out = np.empty((self.cells, 2), np.uint32)
        else:
            w = self.SHAPES[name]
            out = np.empty((self.M, w) if w > 1 else (self.M,), np.float32)
        if self.L.rocl_download(self.h, name.encode(), out.ctypes.data, out.nbytes) != 0:
            raise RuntimeError("rocl_download " + name)
        return out

    def reset_ids(self):
        self.L.rocl_reset_ids(self.h)

    def step(self):
        perm = np.empty(self.M, np.uint32)
        if self.L.rocl_step(self.h, perm.ctypes.data) != 0:
            raise RuntimeError("rocl_step: " + self.L.rocl_last_error(self.h).decode(errors="replace"))
        return perm

    def kernel_times(self, reset=True):
        """{kernel name: (accumulated device microseconds, launches)} from OpenCL profiling events"""
        names, us, n = (C.c_char_p * 32)(), (C.c_double * 32)(), (C.c_uint * 32)()
        k = self.L.rocl_kernel_times(self.h, names, us, n, 32, int(reset))
        return {names[i].decode(): (float(us[i]), int(n[i])) for i in range(k)}

    def close(self):
        if self.h:
            self.L.rocl_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()
