"""CPU-side tests (-m "not gpu"): oracle self-checks against golden vectors, host logic, C-ABI exports."""
import ctypes
import os
import re

import numpy as np

from oracle import oracle_py as O
from realtimeparticles_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rtp_cuda.h")).read()
    declared = set(re.findall(r"RTP_API\s+[\w\s\*]+?\b(rtp_\w+)\s*\(", hdr))
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    L = ctypes.CDLL(_abi.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert L.rtp_abi_version() == 1


def test_param_struct_layouts_match_reference_sizes():
    # BoidsRuleKernelInputs 16 B, TargetKernelInputs 8 B, FluidKernelInputs 44 B, CloudKernelInputs 52 B
    assert ctypes.sizeof(_abi.BoidsParams) == 16 and ctypes.sizeof(_abi.TargetParams) == 8
    assert ctypes.sizeof(_abi.FluidParams) == 44 and ctypes.sizeof(_abi.CloudParams) == 52
    assert _abi.FluidParams.xsphViscosityCoeff.offset == 40 and _abi.CloudParams.windCoeff.offset == 48


def test_baked_constants_match_survey():
    # constants computed with the reference's own Utils::FloatToStr for the 30^3 grid
    w = O.World(O.FLUIDS, 1024, 0)
    assert w.constant("EFFECT_RADIUS") == np.float32(0.3333333433)
    assert w.constant("EFFECT_RADIUS_SQUARED") == np.float32(0.1111111119)
    assert w.constant("POLY6_COEFF") == np.float32(30836.984375)
    assert w.constant("SPIKY_COEFF") == np.float32(3480.7177734375)
    assert _abi.baked_constant(10.0 / 30.0) == np.float32(0.3333333433)
    assert O.lib().orc_baked_constant(ctypes.c_float(1.0 / 3.0)) == _abi.baked_constant(1.0 / 3.0)


def test_generators_product_equals_oracle():
    a = _abi.gen_box_grid((64, 64, 32), (-5, -5, -5), (5, 0, 0))
    b = O.gen_box_grid((64, 64, 32), (-5, -5, -5), (5, 0, 0))
    assert np.array_equal(a, b) and a.shape == (131072, 4)
    assert a[1, 2] == np.float32(-5 + 0.15625) and a[32, 1] == np.float32(-5 + 0.078125)
    a = _abi.gen_sphere_grid((8, 8, 8), (-10 / 6,) * 3, (10 / 6,) * 3)
    b = O.gen_sphere_grid((8, 8, 8), (-10 / 6,) * 3, (10 / 6,) * 3)
    assert np.array_equal(a, b)
    a = _abi.gen_random_box(1000, (-5, -10, -5), (5, -5, 5), 1)
    b = O.gen_random_box(1000, (-5, -10, -5), (5, -5, 5), 1)
    assert np.array_equal(a, b)


def test_planar_generators_of_the_2d_presets():
    # rectangle / circle lattices in a coordinate plane (utils/Geometry.cpp:8-196), used by the 2D presets of the models
    f = np.float32
    for plane, (a, b) in ((_abi.PLANE_XY, (0, 1)), (_abi.PLANE_XZ, (0, 2)), (_abi.PLANE_YZ, (1, 2))):
        res, start, end = (7, 5), (-1.5, 0.25, -3.0), (2.5, 1.75, 0.5)
        r = _abi.gen_rectangle_grid(res, start, end, plane)
        assert r.shape == (35, 4) and np.all(r[:, 3] == 0)
        off = 3 - a - b
        assert np.all(r[:, off] == f(start[off]))  # off-plane coordinate: the start's
        da, db = (f(end[a]) - f(start[a])) / f(res[0]), (f(end[b]) - f(start[b])) / f(res[1])
        ia, ib = np.divmod(np.arange(35), res[1])
        assert np.array_equal(r[:, a], f(start[a]) + ia.astype(f) * da) and np.array_equal(r[:, b], f(start[b]) + ib.astype(f) * db)
        c = _abi.gen_circle_grid(res, start, end, plane)
        ctr = [f(start[k]) + (f(end[k]) - f(start[k])) / f(2) for k in range(3)]
        assert np.all(c[:, off] == ctr[off])
        rad = np.hypot(c[:, a] - ctr[a], c[:, b] - ctr[b])
        half_diag = np.sqrt(sum((end[k] - start[k]) ** 2 for k in range(3))) / 2
        assert np.allclose(rad, (np.arange(35) % res[1] + 1) * half_diag / res[1], rtol=1e-5)
    from oracle import ref_py as R
    if R.available():  # bit for bit against the reference's own Geometry.cpp
        for plane in (0, 1, 2):
            for res, start, end in (((64, 64), (0.0, -5.0, -5.0), (0.0, 0.0, 0.0)), ((32, 16), (0.0, 2.0, -1.0), (0.0, 4.0, 1.0)),
                                    ((256, 512), (0.0, -10 / 6, -10 / 6), (0.0, 10 / 6, 10 / 6))):
                assert np.array_equal(_abi.gen_rectangle_grid(res, start, end, plane), R.generate_2d_grid(0, plane, res, start, end))
                assert np.array_equal(_abi.gen_circle_grid(res, start, end, plane), R.generate_2d_grid(1, plane, res, start, end))
        assert np.array_equal(_abi.gen_random_box(8192, (0.0, -10.0, -5.0), (0.0, 0.0, 5.0), 1),
                              R.generate_2d_grid(0, 2, (128, 64), (0.0, -10.0, -5.0), (0.0, 0.0, 5.0), random=True, seed=1))


def test_generators_match_reference_golden():
    # point sets produced by the reference's own Geometry.cpp (tests/golden/make_golden.py), also where oracle/_ref is absent
    import sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gdir)
    import make_golden
    gold = np.load(os.path.join(gdir, "generators.npz"))
    for name, (product, _) in make_golden.generator_cases().items():
        assert np.array_equal(product(), gold[name]), name


def test_oracle_sort_is_stable():
    rng = np.random.default_rng(0)
    k = rng.integers(0, 500, size=10000).astype(np.uint32)
    ks, perm = O.sort_keys(k)
    assert np.array_equal(ks, np.sort(k, kind="stable"))
    assert np.array_equal(perm, np.argsort(k, kind="stable").astype(np.uint32))


def test_oracle_dam_initial_state_statistics():
    # 6750 of 27000 cells occupied, 19.4 / occupied cell, max 45, cap never binds
    from scenarios import make_fluids
    p = make_fluids(gpu=False, jacobi=1)
    p.w.run_stage("PREDICT_POS")
    p.w.run_stage("FILL_CELL_IDS")
    p.w.run_stage("SORT_BY_CELL")
    p.w.run_stage("BUILD_CELL_TABLE")
    se = p.w.download("START_END_CELL").astype(np.int64)
    cnt = np.where(se[:, 1] >= se[:, 0], se[:, 1] - se[:, 0] + 1, 0)
    assert (cnt > 0).sum() == 6750 and cnt.max() == 45
    keys = p.w.download("CELL_ID")
    assert (np.diff(keys.astype(np.int64)) >= 0).all()
    # quirk 1: the cell holding sorted index 0 keeps start = 1
    assert se[keys[0], 0] == 1


def test_oracle_quirks_table():
    from scenarios import make_fluids
    verts = np.array([[5.0, 0.0, 0.0, 0], [0.0, 5.0, 0.0, 0], [0.0, 0.0, 5.0, 0], [-4.99, -4.99, -4.99, 0]], np.float32)
    p = make_fluids(M=1024, verts=verts, gpu=False, jacobi=1)
    p.w.upload("PRED_POS", p.w.download("POS"))
    p.w.run_stage("FILL_CELL_IDS")
    ids = p.w.download("CELL_ID")[:4]
    # x = +W -> index RES -> id >= C (ignored); y/z = +W alias into the next row/column (quirk 5)
    assert ids[0] == 30 * 900 + 15 * 30 + 15 and ids[0] >= 27000
    assert ids[1] == 15 * 900 + 30 * 30 + 15
    assert ids[2] == 15 * 900 + 15 * 30 + 30
    assert ids[3] == 0
    # tail keys 2C + i
    assert p.w.download("CELL_ID")[4] == 54000 + 4


def test_model_json_round_trip_and_errors():
    from realtimeparticles_b200 import models
    js = models.initFluidsJson
    assert js["Fluids"]["Nb Jacobi Iterations"] == [2, 1, 6]
    merged = models._merge_patch({"a": {"b": 1, "c": 2}}, {"a": {"b": None, "d": 3}})
    assert merged == {"a": {"c": 2, "d": 3}}
    assert list(models.initCloudsJson["Clouds"].keys())[0] == "Enable Temperature Smoothing"


def test_no_gpu_means_loud_failure_not_fallback():
    if _abi.lib().rtp_device_count() > 0:
        return
    try:
        _abi.Handle(_abi.FLUIDS, 1024, 1024)
    except _abi.RtpError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("rtp_create must fail without a CUDA device")


def test_cpp_dropin_library_exports():
    # Physics::CUDA::{Boids,Fluids,Clouds} behind the unmodified Physics::Model (built only where /root/reference exists)
    lib = os.path.join(ROOT, "realtimeparticles_b200", "lib", "librtp_models.so")
    if not os.path.exists(lib):
        import pytest
        pytest.skip("librtp_models.so not built (needs /root/reference)")
    L = ctypes.CDLL(lib)
    for name in ("rtpm_create", "rtpm_destroy", "rtpm_update", "rtpm_reset", "rtpm_update_input_json", "rtpm_get_input_json",
                 "rtpm_handle", "rtpm_nb_particles", "rtpm_is_init", "rtpm_pause", "rtpm_set_boundary", "rtpm_set_step_flags"):
        assert hasattr(L, name), name


def test_oracle_boids_config1_loses_no_boid():
    # BASELINE config 1 on the oracle: with fast_normalize(0) = 0 (OpenCL 1.2 s6.12.5) no boid turns NaN over 300 steps
    # (the NaN-propagating variant had lost 9 % of them by step 200); the CUDA path runs the full 1000 steps bit-exact
    # against this in tests/test_gpu_parity.py
    from scenarios import make_boids
    p = make_boids(M=131072, N=512, gpu=False)
    for _ in range(300):
        p.w.step(O.STEP_PHYSICS)
    pos, vel = p.w.download("POS")[:512], p.w.download("VEL")[:512]
    assert np.isfinite(pos).all() and np.isfinite(vel).all()
    assert (np.abs(pos[:, :3] + 5.0).max(axis=1) < 1e-6).sum() == 0


def test_oracle_fast_normalize_zero_vector():
    # one boid pair placed symmetrically so that a rule sum cancels exactly: acceleration stays finite
    verts = np.array([[0.1, 0.0, 0.0, 0.0], [-0.1, 0.0, 0.0, 0.0], [0.0, 0.1, 0.0, 0.0]], np.float32)
    w = O.World(O.BOIDS, 1024, 3)
    pos = np.full((1024, 4), np.inf, np.float32)
    pos[:, 3] = 0
    pos[:3] = verts
    w.upload("POS", pos)
    w.upload("VEL", np.zeros((1024, 4), np.float32))  # fast_normalize(velocity[e]) of zero velocities
    w.reset_ids()
    w.step(O.STEP_PHYSICS)
    assert np.isfinite(w.download("ACC")[:3]).all() and np.isfinite(w.download("VEL")[:3]).all()


def test_target_trajectory_matches_reference():
    # rtp_target_* (host code of the library) against Physics::Target of the reference itself: golden positions generated
    # through oracle/_ref (tests/golden/make_golden.py), and live when oracle/_ref is built
    gold = np.load(os.path.join(ROOT, "tests", "golden", "target_trajectory.npz"))
    for key, dim, vel in (("dim3_v05", 3, 0.5), ("dim2_v20", 2, 2.0)):
        t = _abi.Target(10)
        got = np.array([t.update(dim, vel) for _ in range(400)], np.float32)
        assert np.array_equal(got, gold[key]), key
        assert np.abs(got).max() <= 4.8 + 1e-5 and (dim == 3 or np.all(got[:, 0] == 0.0))
    from oracle import ref_py as R
    if R.available():
        assert np.array_equal(R.target_trajectory(10, 3, 0.5, 50), gold["dim3_v05"][:50])


def test_bench_reference_arm_contract():
    # `bench.py --impl reference`: one JSON line with the GPU arm's metric / unit / config, impl = "reference", a cpu_baseline
    # describing this run and an e2e object without transfers; under torchrun only rank 0 prints
    import json
    import subprocess
    import sys
    sys.path.insert(0, ROOT)
    import bench
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "particle-updates/s" and d["higher_is_better"] is True
    assert d["config"] == bench.headline_config(1) and d["config"]["workload"] == "pbf_dam_130k_I3_vorticity_xsph"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ""  # the other ranks exit 0 without work
