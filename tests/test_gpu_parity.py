"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): cell ids, sorted permutation, cell start/end table BIT-EXACT; single-step positions
and velocities within 1e-5 relative; 100-step aggregate invariants within 1 %.
What is enforced here is stricter: the CUDA kernels implement the same canonical fp32 operation sequence as the
oracle (DESIGN.md "Canonical arithmetic") and sum in the reference's order, so EVERY field is required to be
bit-identical (TOL = 0), single-step and multi-step. rel_err() is only used to print how far off a failure is.
"""
import numpy as np
import pytest

from oracle import oracle_py as O
from realtimeparticles_b200 import _abi

from scenarios import (M130K, make_boids, make_clouds, make_fluids, pbf_invariants, rel_err)

pytestmark = pytest.mark.gpu

TOL = 0.0  # bit-exact; the north_star bar is 1e-5 relative


def assert_ids_exact(p, N=None):
    for f in ("CELL_ID", "PERM", "START_END_CELL"):
        a, b = p.get(f)
        assert np.array_equal(a, b), "%s not bit-exact: %d mismatches" % (f, int((a != b).sum()))


def assert_close(p, fields, tol=TOL, N=None):
    N = p.N if N is None else N
    for f in fields:
        a, b = p.get(f)
        a, b = a[:N], b[:N]
        if tol == 0.0:
            same = (a.view(np.uint32) == b.view(np.uint32)) | ((a == b) & np.isfinite(a)) | (np.isnan(a) & np.isnan(b))
            assert same.all(), "%s not bit-exact: %d of %d elements differ, rel err %.3g" % (
                f, int((~same).sum()), same.size, rel_err(b, a))
        else:
            e = rel_err(b, a)
            assert e <= tol, "%s rel err %.3g > %.1g" % (f, e, tol)


# ---------------------------------------------------------------- exact math helpers

def test_inrange_sqrt_rcp_are_ieee_exact():
    # every float in [1e-17, 1e5]: the kernels' range-check-free sqrt / reciprocal == __fsqrt_rn / __frcp_rn
    h = _abi.Handle(_abi.BOIDS, 1024, 0)
    bad_sqrt, bad_rcp = h.selftest_math(1e-17, 1e5)
    assert bad_sqrt == 0 and bad_rcp == 0, (bad_sqrt, bad_rcp)


# ---------------------------------------------------------------- sort

@pytest.mark.parametrize("n,bits,hi", [(0, 8, 200), (1, 8, 200), (31, 16, 60000), (1000, 16, 55000), (4097, 32, 2**32 - 1),
                                        (131072, 16, 54000), (131072, 20, 10**6), (1 << 20, 22, 3456000),
                                        ((1 << 21) + 3, 32, 2**32 - 1)])
def test_sort_matches_oracle(n, bits, hi):
    rng = np.random.default_rng(1234 + n)
    keys = rng.integers(0, hi + 1, size=n, dtype=np.uint64).astype(np.uint32)
    h = _abi.Handle(_abi.BOIDS, 1024, 0)
    ks, perm = h.sort_keys_host(keys, bits)
    ko, po = O.sort_keys(keys)
    assert np.array_equal(ks, ko)
    assert np.array_equal(perm, po)  # stability: equal keys keep their input order


def test_sort_heavy_duplicates_is_stable():
    # all particles in 3 cells: every tile publishes huge same-digit runs
    rng = np.random.default_rng(7)
    keys = rng.choice(np.array([5, 6, 26999], np.uint32), size=200000)
    h = _abi.Handle(_abi.BOIDS, 1024, 0)
    ks, perm = h.sort_keys_host(keys, 16)
    assert np.array_equal(ks, np.sort(keys, kind="stable"))
    assert np.array_equal(perm, np.argsort(keys, kind="stable").astype(np.uint32))


# ---------------------------------------------------------------- fluids

def test_fluids_dam_130k_single_step():
    p = make_fluids(jacobi=3)
    p.step(O.STEP_PHYSICS | O.STEP_DEBUG_FIELDS)
    assert_ids_exact(p)
    assert_close(p, ("POS", "VEL", "PRED_POS"))
    assert_close(p, ("DENSITY", "CONST_FACTOR"))
    assert_close(p, ("CORR_POS", "VORT"))


@pytest.mark.parametrize("jacobi,vort,art", [(1, 1, 1), (2, 0, 1), (6, 1, 0)])
def test_fluids_variants(jacobi, vort, art):
    p = make_fluids(M=16384, res=(32, 32, 16), jacobi=jacobi, isVorticityConfEnabled=vort, isArtPressureEnabled=art)
    for _ in range(2):
        p.step(O.STEP_PHYSICS)
    assert_close(p, ("POS", "VEL", "PRED_POS", "DENSITY"))


def test_fluids_ragged_and_tail():
    # N < M: the inactive tail keeps keys 2C+i, +inf positions and the identity permutation
    verts = O.gen_box_grid((16, 16, 16), (-2.0, -2.0, -2.0), (2.0, 2.0, 2.0))
    p = make_fluids(M=8192, N=3000, verts=verts, jacobi=2)
    p.step(O.STEP_PHYSICS)
    assert_ids_exact(p)
    assert_close(p, ("POS", "VEL"), N=3000)
    pos = p.h.download("p_pos")
    assert np.isinf(pos[3000:, :3]).all()


@pytest.mark.parametrize("n", [0, 1, 2])
def test_fluids_tiny_counts(n):
    verts = np.array([[0.1, 0.2, 0.3, 0.0], [0.15, 0.2, 0.3, 0.0]], np.float32)
    p = make_fluids(M=1024, N=n, verts=verts, jacobi=2)
    p.step(O.STEP_PHYSICS)
    assert_ids_exact(p)
    if n:
        assert_close(p, ("POS", "VEL"), N=n)


def test_fluids_wall_quirks():
    # particles exactly on the +W walls (cell index == RES) and outside the box
    rng = np.random.default_rng(3)
    verts = rng.uniform(-5.2, 5.2, size=(4096, 4)).astype(np.float32)
    verts[:, 3] = 0
    verts[:64, 0] = 5.0
    verts[64:128, 1] = 5.0
    verts[128:192, 2] = 5.0
    verts[192:200] = [5.0, 5.0, 5.0, 0.0]
    p = make_fluids(M=4096, verts=verts, jacobi=2)
    p.step(O.STEP_PHYSICS)
    assert_ids_exact(p)
    assert_close(p, ("POS", "VEL"))


def test_fluids_cap_binds():
    # 400 particles in one cell with cap 100: adjustEndCell keeps MAX+1 (quirk 3)
    rng = np.random.default_rng(5)
    verts = np.zeros((2048, 4), np.float32)
    verts[:, :3] = rng.uniform(0.01, 0.32, size=(2048, 3))
    verts[400:, :3] = rng.uniform(-4.9, 4.9, size=(1648, 3))
    p = make_fluids(M=2048, verts=verts, jacobi=1)
    p.step(O.STEP_PHYSICS)
    assert_ids_exact(p)
    se = p.h.download("c_startEndPartID")
    cnt = se[:, 1].astype(np.int64) - se[:, 0] + 1
    assert cnt.max() == 101


def test_fluids_100_step_invariants():
    # 16k-particle dam: mean density error and kinetic energy within 1 % after 100 steps
    p = make_fluids(M=16384, res=(32, 32, 16), jacobi=3)
    for _ in range(100):
        p.step(O.STEP_PHYSICS)
    (do, dg), (vo, vg) = p.get("DENSITY"), p.get("VEL")
    eo, ko = pbf_invariants(do, vo, p.N)
    eg, kg = pbf_invariants(dg, vg, p.N)
    assert abs(eg - eo) <= 0.01 * eo, (eg, eo)  # north_star bar
    assert abs(kg - ko) <= 0.01 * ko, (kg, ko)
    assert_close(p, ("POS", "VEL", "DENSITY"))  # and in fact bit-identical after 100 steps


def test_lists_on_off_bit_identical(monkeypatch):
    # the neighbour-list engine (sweep.cuh) must not change a single bit relative to the plain 27-cell traversal,
    # also when the margin is so small that the lists are invalidated and rebuilt inside the step
    outs = []
    for env in ({"RTP_NBR_LISTS": "0"}, {"RTP_NBR_LISTS": "1"}, {"RTP_NBR_LISTS": "1", "RTP_NBR_MARGIN": "0.02"},
                {"RTP_NBR_LISTS": "1", "RTP_NBR_CAP": "64", "RTP_HIT_CAP": "48"}, {"RTP_NBR_LISTS": "1", "RTP_TILED_BUILD": "0"},
                {"RTP_NBR_LISTS": "1", "RTP_TILED_BUILD": "0", "RTP_NBR_MARGIN": "0.02"}):
        for k in ("RTP_NBR_LISTS", "RTP_NBR_MARGIN", "RTP_NBR_CAP", "RTP_HIT_CAP", "RTP_TILED_BUILD"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        p = make_fluids(M=16384, res=(32, 32, 16), jacobi=3)
        for _ in range(10):
            p.step(O.STEP_PHYSICS, oracle=False)
        outs.append([p.h.download(f) for f in ("p_pos", "p_vel", "p_density", "p_vort")])
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)


def test_stragglers_occur_and_stay_exact(monkeypatch):
    # particles that change cell between two solver iterations cannot use their margin list; they are queued by the
    # correction kernel and swept by whole warps (sweep.cuh STRAGGLERS). The collapsing dam must produce such particles,
    # and the run must stay bit-identical to the list-free traversal and to the oracle.
    for k in ("RTP_NBR_LISTS", "RTP_NBR_MARGIN", "RTP_NBR_CAP", "RTP_HIT_CAP"):
        monkeypatch.delenv(k, raising=False)
    p = make_fluids(M=16384, res=(32, 32, 16), jacobi=3)
    changed = 0
    for _ in range(30):
        p.step(O.STEP_PHYSICS)
        st = p.h.list_stats()
        assert st["particles"] == 16384 and st["margin_overflow"] == 0 and st["moved_beyond_bound"] == 0
        changed += st["cell_changed"]
    assert changed > 0
    monkeypatch.setenv("RTP_NBR_LISTS", "0")
    q = make_fluids(M=16384, res=(32, 32, 16), jacobi=3)
    for _ in range(30):
        q.step(O.STEP_PHYSICS, oracle=False)
    for f in ("p_pos", "p_vel", "p_density", "p_vort"):
        assert np.array_equal(p.h.download(f), q.h.download(f)), f
    for f in ("POS", "VEL", "DENSITY", "VORT"):
        a, b = p.get(f)
        assert np.array_equal(a, b), f  # 30 steps, still bit-identical with the oracle
    assert q.h.list_stats()["particles"] == 0  # lists disabled: nothing to report


def test_list_allocation_failure_falls_back_to_plain_traversal(monkeypatch):
    # the lists cost 1.66 kB per particle of max_particles; when they cannot be allocated rtp_create must not fail: the
    # sweeps take the plain 27-cell traversal (bit-identical)
    monkeypatch.setenv("RTP_TEST_FAIL_LIST_ALLOC", "1")
    p = make_fluids(M=16384, res=(32, 32, 16), jacobi=3)
    monkeypatch.delenv("RTP_TEST_FAIL_LIST_ALLOC")
    q = make_fluids(M=16384, res=(32, 32, 16), jacobi=3)
    for _ in range(5):
        p.step(O.STEP_PHYSICS, oracle=False)
        q.step(O.STEP_PHYSICS, oracle=False)
    assert p.h.list_stats()["particles"] == 0 and q.h.list_stats()["particles"] == 16384
    for f in ("p_pos", "p_vel", "p_density"):
        assert np.array_equal(p.h.download(f), q.h.download(f)), f


def test_fluid_params_are_validated():
    # artPressureExp outside [1, 6] or a radius factor outside (0, 1) would silently change the arithmetic (the exponent
    # is an unrolled product; radius >= 1 divides by poly6 = 0): rejected like the reference's UI ranges (Fluids.cpp:52-74)
    h = _abi.Handle(_abi.FLUIDS, 1024, 0)
    for exp, radius in ((0, 0.006), (7, 0.006), (4, 1.0), (4, 0.0)):
        with pytest.raises(_abi.RtpError):
            h.set_fluid_params(_abi.FluidParams(450.0, 600.0, 0.010, 3, 1, radius, 0.001, exp, 1, 0.0004, 0.0001), 3)
    h.set_fluid_params(_abi.FluidParams(450.0, 600.0, 0.010, 3, 0, 1.0, 0.001, 9, 1, 0.0004, 0.0001), 3)  # disabled: not looked at
    h.set_fluid_params(_abi.FluidParams(450.0, 600.0, 0.010, 3, 1, 0.015, 0.001, 6, 1, 0.0004, 0.0001), 3)


@pytest.mark.parametrize("case", ["FLUIDS_BOMB", "FLUIDS_DROP"])
def test_fluid_presets_bomb_and_drop_vs_oracle(case):
    # the remaining fluids presets (Fluids.cpp:339-380): 65k sphere "bomb"; 4k drop over a 65k floor layer
    from realtimeparticles_b200 import models
    M = 131072
    py = models.CreateModel(1, models.ModelParams(currNbParticles=M, maxNbParticles=M, boxSize=(10, 10, 10), gridRes=(30, 30, 30),
                                                  pCase=getattr(models.PhysicsCase, case), dimension=models.Dimension.dim3D))
    N = py.nbParticles()
    assert N == (65536 if case == "FLUIDS_BOMB" else 4096 + 65536)
    verts = py.download("p_pos")[:N]
    p = make_fluids(M=M, N=N, verts=verts, jacobi=2)
    for _ in range(3):
        p.step(O.STEP_PHYSICS)
    assert_ids_exact(p)
    assert_close(p, ("POS", "VEL", "PRED_POS", "DENSITY", "VORT"), N=N)


def test_gl_interop_without_a_gl_context_fails_cleanly():
    # rtp_register_gl (cudaGraphicsGLRegisterBuffer on the VBO names of ModelParams, Model.hpp:67-70) needs a current OpenGL
    # context; the GPU box is headless. The entry points must exist, refuse with RTP_ERR_CUDA + a message, and leave the
    # handle fully usable (the zero-copy path itself needs a display: INTEGRATION.md section 4).
    p = make_fluids(M=4096, res=(16, 16, 16), jacobi=2)
    for field in ("p_pos", "p_col"):
        with pytest.raises(_abi.RtpError) as ei:
            p.h.register_gl(field, 1)
        assert "cudaGraphicsGLRegisterBuffer" in str(ei.value)
    with pytest.raises(_abi.RtpError):
        p.h.register_gl("p_vel", 1)  # not a shared buffer
    p.step(O.STEP_PHYSICS | O.STEP_RENDER_AUX)
    assert_close(p, ("POS", "VEL", "COL"))


def test_step_n_graph_replay_equals_single_steps():
    a = make_fluids(M=16384, res=(32, 32, 16), jacobi=3)
    b = make_fluids(M=16384, res=(32, 32, 16), jacobi=3)
    for _ in range(5):
        a.h.step(_abi.STEP_PHYSICS)
    b.h.step_n(5, _abi.STEP_PHYSICS)
    a.h.sync()
    b.h.sync()
    for f in ("p_pos", "p_vel", "p_cellID", "RadixSortIndices"):
        assert np.array_equal(a.h.download(f), b.h.download(f)), f


def test_full_update_with_camera_sort():
    # update() == physics + render-side kernels + camera sort (Fluids.cpp:400-471); the camera sort reorders state
    p = make_fluids(M=16384, res=(32, 32, 16), jacobi=2)
    flags = O.STEP_PHYSICS | O.STEP_RENDER_AUX | O.STEP_CAMERA_SORT
    p.step(flags)
    for f in ("CAMERA_DIST", "CAMERA_PERM", "PART_DETECTOR"):
        a, b = p.get(f)
        assert np.array_equal(a, b), f
    assert_close(p, ("POS", "VEL", "COL"))
    p.step(flags)
    assert_close(p, ("POS", "VEL"))


# ---------------------------------------------------------------- boids

def test_boids_512_config1_1000_steps_bit_exact():
    # BASELINE config 1: 512 boids in M = 131072 (exercises the tail), the full 1000 headless steps; the whole boids step is
    # bit-exact by construction. fast_normalize(0) = 0 (OpenCL 1.2 s6.12.5): no boid may turn NaN or get parked in a corner
    # (boids.cl:120-124, :258).
    p = make_boids(M=M130K, N=512)
    for s in range(1000):
        p.step(O.STEP_PHYSICS)
        if s in (0, 49, 199, 999):
            assert_ids_exact(p)
            assert_close(p, ("POS", "VEL", "ACC"), N=512)
    pos, vel = p.h.download("p_pos")[:512], p.h.download("p_vel")[:512]
    assert np.isfinite(pos).all() and np.isfinite(vel).all()
    speed = np.linalg.norm(vel[:, :3], axis=1)
    assert speed.min() > 0.09 and speed.max() < 0.51  # bd_updateVel clamps the speed to [0.2, 1] * velocityScale
    assert (np.abs(pos[:, :3] + 5.0).max(axis=1) < 1e-6).sum() == 0  # nobody parked at (-5, -5, -5)


def test_boids_130k_single_step():
    p = make_boids(M=M130K, N=M130K, res=(64, 64, 32))
    p.step(O.STEP_PHYSICS)
    assert_ids_exact(p)
    assert_close(p, ("POS", "VEL", "ACC"))


def test_boids_periodic_and_target():
    p = make_boids(M=4096, N=4096, res=(16, 16, 16))
    p.both(lambda w: w.set_boundary(1), lambda h: h.set_boundary(1))
    bo, ba = O.default_boids_params(), _abi.BoidsParams(0.5, 1.6, 1.6, 1.45)
    p.w.set_boids_params(bo, O.TargetParams(2.0, -1), (0.5, 0.25, -0.5, 0.0), True)
    p.h.set_boids_params(ba, _abi.TargetParams(2.0, -1), (0.5, 0.25, -0.5, 0.0), True)
    for _ in range(30):
        p.step(O.STEP_PHYSICS)
    assert_ids_exact(p)
    assert_close(p, ("POS", "VEL", "ACC"))


def test_boids_2d():
    rng = np.random.default_rng(11)
    verts = np.zeros((2048, 4), np.float32)
    verts[:, 1:3] = rng.uniform(-1.6, 1.6, size=(2048, 2))
    p = make_boids(M=2048, N=2048, dim=2, verts=verts)
    for _ in range(5):
        p.step(O.STEP_PHYSICS)
    assert_ids_exact(p)
    assert_close(p, ("POS", "VEL", "ACC"))


# ---------------------------------------------------------------- clouds

def test_clouds_init_fields():
    p = make_clouds(M=65536, N=65536)
    for f in ("TEMP", "VAPOR_DENS"):
        assert np.array_equal(p.init_gpu[f], p.w.download(f)), (f, rel_err(p.init_gpu[f], p.w.download(f)))


def test_clouds_130k_single_step():
    p = make_clouds()
    p.step(O.STEP_PHYSICS)
    assert_ids_exact(p)
    assert_close(p, ("POS", "VEL", "PRED_POS", "TOT_CORR_POS"))
    assert_close(p, ("TEMP", "VAPOR_DENS", "CLOUD_DENS", "BUOYANCY", "PART_ID", "DENSITY", "CONST_FACTOR"))
    assert_close(p, ("LAPLACIAN_TEMP", "CONST_FACTOR_TEMP", "CORR_TEMP", "CLOUD_GEN", "VORT"))


def test_clouds_periodic_images_and_steps():
    # homogeneous fill of the whole box so that the x/z periodic images and the y walls are all exercised
    p = make_clouds(M=32768, N=32768, region=((-5.0, -10.0, -5.0), (5.0, 10.0, 5.0)), jacobi=2)
    for _ in range(3):
        p.step(O.STEP_PHYSICS | O.STEP_RENDER_AUX | O.STEP_CAMERA_SORT)
    assert_close(p, ("POS", "VEL", "TEMP", "VAPOR_DENS", "CLOUD_DENS", "COL"))


def test_clouds_no_smoothing_no_vorticity():
    p = make_clouds(M=16384, N=16384, isTempSmoothingEnabled=0)
    fo, fa = O.default_fluid_params(), _abi.FluidParams(450.0, 600.0, 0.010, 3, 1, 0.006, 0.001, 4, 0, 0.0004, 0.0001)
    fo.isVorticityConfEnabled = 0
    p.w.set_fluid_params(fo, 2)
    p.h.set_fluid_params(fa, 2)
    p.step(O.STEP_PHYSICS)
    assert_ids_exact(p)
    assert_close(p, ("POS", "VEL", "TEMP"))


# ---------------------------------------------------------------- full-size properties

def test_pbf_2m_properties():
    # size-independent properties at a size the oracle would not finish quickly: sortedness, permutation validity,
    # table <-> keys consistency, finite state (SURVEY 8d config 5 geometry scaled to 2M: box 40x20x20)
    N = 1 << 21
    verts = _abi.gen_box_grid((256, 128, 64), (-20.0, -10.0, -10.0), (20.0, 0.0, 0.0))
    h = _abi.Handle(_abi.FLUIDS, N, N, (40, 20, 20), (120, 60, 60))
    h.set_fluid_params(_abi.FluidParams(450.0, 600.0, 0.010, 3, 1, 0.006, 0.001, 4, 1, 0.0004, 0.0001), 3)
    h.upload("p_pos", verts)
    h.reset_ids()
    h.step_n(2, _abi.STEP_PHYSICS)
    h.sync()
    keys, perm, se = h.download("p_cellID"), h.download("RadixSortIndices"), h.download("c_startEndPartID")
    assert (np.diff(keys.astype(np.int64)) >= 0).all()
    assert np.array_equal(np.sort(perm), np.arange(N, dtype=np.uint32))
    occ = se[:, 1] >= se[:, 0]
    cells = np.nonzero(occ)[0]
    assert np.array_equal(keys[se[cells, 0]], cells.astype(np.uint32)) or se[cells[0], 0] == 1
    assert np.array_equal(keys[se[cells, 1]], cells.astype(np.uint32))
    pos, vel = h.download("p_pos"), h.download("p_vel")
    assert np.isfinite(pos).all() and np.isfinite(vel).all()
    d = h.download("p_density")
    assert 0.0 < d.mean() < 2000.0


# ---------------------------------------------------------------- the C++ drop-in (Physics::CUDA::* : Physics::Model)

def _cpp_models():
    import ctypes as C
    import os
    lib = os.path.join(os.path.dirname(_abi.LIB_PATH), "librtp_models.so")
    if not os.path.exists(lib):
        pytest.skip("librtp_models.so not built (needs /root/reference at build time)")
    L = C.CDLL(lib)
    L.rtpm_create.restype = C.c_void_p
    L.rtpm_create.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.rtpm_handle.restype = C.c_void_p
    L.rtpm_handle.argtypes = [C.c_void_p, C.c_int]
    for f in ("rtpm_destroy", "rtpm_update", "rtpm_reset"):
        getattr(L, f).argtypes = [C.c_void_p]
    L.rtpm_nb_particles.argtypes = [C.c_void_p]
    L.rtpm_nb_particles.restype = C.c_uint64
    L.rtpm_is_init.argtypes = [C.c_void_p]
    L.rtpm_update_input_json.argtypes = [C.c_void_p, C.c_char_p]
    L.rtpm_get_input_json.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.rtpm_set_step_flags.argtypes = [C.c_void_p, C.c_int, C.c_uint]
    return L, C


@pytest.mark.parametrize("kind", ["fluids", "boids", "clouds", "fluids2d", "boids2d", "fluids_bomb", "fluids_drop",
                                  "clouds_homogeneous", "boids_target", "boids_target2d"])
def test_cpp_model_equals_python_model(kind):
    """Physics::CUDA::X driven through the reference's own Model interface == the Python mirror, bit for bit.
    Covers every preset (Fluids.cpp:339-380 Dam / Bomb / Drop, Clouds.cpp:449-471 Cumulus / Homogeneous, Boids.cpp:235-257)
    and the boids target rule with its Perlin-noise trajectory (Boids.cpp:351-358: the C++ host runs the reference's own
    Physics::Target, the Python host rtp_target_update)."""
    from realtimeparticles_b200 import models
    L, C = _cpp_models()
    variant = kind
    dim3 = not kind.endswith("2d")  # 2D: the presets live in the YZ plane (Generate2DGrid), same kernels
    kind = kind[:-2] if not dim3 else kind
    target = kind == "boids_target"
    PC = models.PhysicsCase
    t, case, box, grid = {"fluids": (1, PC.FLUIDS_DAM, (10, 10, 10), (30, 30, 30)),
                          "fluids_bomb": (1, PC.FLUIDS_BOMB, (10, 10, 10), (30, 30, 30)),
                          "fluids_drop": (1, PC.FLUIDS_DROP, (10, 10, 10), (30, 30, 30)),
                          "boids": (0, PC.BOIDS_LARGE, (10, 10, 10), (30, 30, 30)),
                          "boids_target": (0, PC.BOIDS_MEDIUM, (10, 10, 10), (30, 30, 30)),
                          "clouds": (2, PC.CLOUDS_CUMULUS, (10, 20, 10), (30, 60, 30)),
                          "clouds_homogeneous": (2, PC.CLOUDS_HOMOGENEOUS, (10, 20, 10), (30, 60, 30))}[kind]
    kind = kind.split("_")[0]
    M = 131072
    m = L.rtpm_create(t, M, int(case), 1 if dim3 else 0, (C.c_uint32 * 3)(*box), (C.c_uint32 * 3)(*grid))
    assert m and L.rtpm_is_init(m)
    if kind != "boids":
        assert L.rtpm_update_input_json(m, b'{"Fluids": {"Nb Jacobi Iterations": [3, 1, 6]}}') == 0
        assert L.rtpm_update_input_json(m, b'{"Fluids": {"Nb Jacobi Iterations": "oops"}}') == -1  # "Wrong Json parsing"
        assert L.rtpm_update_input_json(m, b'{"Fluids": {"Nb Jacobi Iterations": [3, 1, 6]}}') == 0
    py = models.CreateModel(t, models.ModelParams(currNbParticles=M, maxNbParticles=M, boxSize=box, gridRes=grid, pCase=case,
                                                  dimension=models.Dimension.dim3D if dim3 else models.Dimension.dim2D))
    if not dim3 and kind == "fluids":
        assert py.nbParticles() == 4096 and np.all(py.download("p_pos")[:4096, 0] == 0.0)
    if kind == "clouds":
        # the reference seeds nothing: both sides draw from glibc rand(); make the two draws identical
        region = ((-5.0, -10.0, -5.0), (5.0, -5.0, 5.0)) if case == PC.CLOUDS_CUMULUS else ((-5.0, -10.0, -5.0), (5.0, 10.0, 5.0))
        verts = _abi.gen_random_box(65536, region[0], region[1], 1)
        py.loadCloudsState(verts)
        h = _abi.Handle.__new__(_abi.Handle)  # borrowed view of the C++ model's handle: never let it destroy it
        h.L, h.h, h.model, h.M, h.N = _abi.lib(), C.c_void_p(L.rtpm_handle(m, t)), t, M, 65536
        pos = np.full((M, 4), np.inf, np.float32)
        pos[:, 3] = 0
        pos[:65536] = verts
        h.upload("p_pos", pos)
        h.init_clouds_fields()
        h.h = None
    if kind != "boids":
        js = py.getInputJson()
        js["Fluids"]["Nb Jacobi Iterations"][0] = 3
        py.updateInputJson(js)
    if target:
        tj = b'{"Boids": {"Target": {"Enable##Target": true, "Radius": [3.0, 1.0, 20.0], "Attract": false}}}'
        assert L.rtpm_update_input_json(m, tj) == 0
        js = py.getInputJson()
        js["Boids"]["Target"].update({"Enable##Target": True, "Radius": [3.0, 1.0, 20.0], "Attract": False})
        py.updateInputJson(js)
        assert py.isTargetActivated()
    assert L.rtpm_nb_particles(m) == py.nbParticles()
    expect = {"fluids_bomb": 65536, "fluids_drop": 4096 + 65536, "clouds_homogeneous": 65536, "boids_target": 16384}
    if variant in expect:
        assert py.nbParticles() == expect[variant], py.nbParticles()
    for _ in range(12 if target else 3):
        L.rtpm_update(m)
        py.update()
    if target:
        assert np.abs(np.array(py.targetPos())).max() > 0.1  # the target has left the origin
    h = _abi.Handle.__new__(_abi.Handle)
    h.L, h.h, h.model, h.M, h.N = _abi.lib(), C.c_void_p(L.rtpm_handle(m, t)), t, M, py.nbParticles()
    h.sync()
    py.sync()
    for f in ("p_pos", "p_vel", "p_col", "p_cellID", "p_cameraDist"):
        a, b = h.download(f), py.download(f)
        assert np.array_equal(a, b, equal_nan=True), f
    h.h = None
    L.rtpm_destroy(m)


@pytest.mark.gpu
def test_stream_ordered_copies_equal_blocking_copies():
    # rtp_upload_async / rtp_download_async from page-locked buffers around the cached-graph update() (the e2e loop of
    # bench.py: one synchronisation per frame) against the blocking calls around plain rtp_step launches: same bits
    import torch
    from realtimeparticles_b200 import models
    n = 131072
    params = models.ModelParams(currNbParticles=n, maxNbParticles=n, boxSize=(10, 10, 10), gridRes=(30, 30, 30),
                                pCase=models.PhysicsCase.FLUIDS_DROP)  # 4k-particle block falling into a 65k-particle pool
    out = []
    for streamed in (False, True):
        m = models.CreateModel(models.ModelType.FLUIDS, params)
        m.setStepFlags(_abi.STEP_PHYSICS)
        pos = torch.from_numpy(m.download("p_pos")).pin_memory().numpy()
        vel = torch.from_numpy(m.download("p_vel")).pin_memory().numpy()
        for _ in range(6):
            if streamed:
                m.upload("p_pos", pos, blocking=False)
                m.upload("p_vel", vel, blocking=False)
                m.update()
                m.download("p_pos", out=pos, blocking=False)
                m.download("p_vel", out=vel, blocking=False)
                m.sync()
            else:
                m.upload("p_pos", pos)
                m.upload("p_vel", vel)
                m.update(replay_graph=False)
                m.download("p_pos", out=pos)
                m.download("p_vel", out=vel)
        out.append((pos.copy(), vel.copy()))
    assert np.array_equal(out[0][0], out[1][0], equal_nan=True) and np.array_equal(out[0][1], out[1][1], equal_nan=True)
    assert np.isfinite(out[0][0][:m.nbParticles()]).all() and np.abs(out[0][1]).max() > 0
