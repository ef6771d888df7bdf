"""Generate tests/golden/*.npz from the REFERENCE'S OWN kernels (oracle/_ref, built from /root/reference by
oracle/ref/Makefile). Run in a container that has /root/reference:   python tests/golden/make_golden.py
The fixtures are small on purpose; every array is an output of the reference's .cl source executed on the CPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O, ref_py as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cases():
    """name -> (world factory taking the World class, number of steps, flags, fields to store)"""
    def fluids(cls):
        M = 2048
        w = cls(O.FLUIDS, M, M)
        w.set_fluid_params(O.default_fluid_params(), 3)
        w.upload("POS", R.generate_3d_grid(0, (16, 16, 8), (-5.0, -5.0, -5.0), (-2.5, -3.75, -3.75)))
        w.upload("VEL", np.zeros((M, 4), np.float32))
        w.reset_ids()
        return w

    def boids(cls):
        M, N = 1024, 512
        w = cls(O.BOIDS, M, N)
        pos = np.full((M, 4), np.inf, np.float32)
        pos[:, 3] = 0
        pos[:N] = R.generate_3d_grid(1, (8, 8, 8), (-10 / 6.0,) * 3, (10 / 6.0,) * 3)
        w.upload("POS", pos)
        w.upload("VEL", pos)
        w.reset_ids()
        return w

    def clouds(cls):
        M = 4096
        w = cls(O.CLOUDS, M, M, (10, 20, 10), (30, 60, 30))
        w.set_fluid_params(O.default_fluid_params(), 2)
        w.set_cloud_params(O.default_cloud_params())
        w.upload("POS", R.generate_3d_grid(0, (16, 16, 16), (-5.0, -10.0, -5.0), (-2.0, -8.0, -2.0), random=True, seed=1))
        w.upload("VEL", np.zeros((M, 4), np.float32))
        w.upload("CLOUD_DENS", np.zeros(M, np.float32))
        w.upload("PART_ID", np.arange(M, dtype=np.float32))
        w.init_clouds_fields()
        w.reset_ids()
        return w

    phys, full = O.STEP_PHYSICS, O.STEP_PHYSICS | O.STEP_RENDER_AUX | O.STEP_CAMERA_SORT
    ids = ["CELL_ID", "PERM", "START_END_CELL"]
    return {
        "fluids_2k_1step": (fluids, 1, phys, ids + ["POS", "VEL", "PRED_POS", "DENSITY", "CONST_FACTOR", "CORR_POS", "VORT"]),
        "fluids_2k_3steps_full_update": (fluids, 3, full, ["CAMERA_DIST", "CAMERA_PERM", "POS", "VEL", "COL"]),
        "boids_512_5steps": (boids, 5, phys, ids + ["POS", "VEL", "ACC"]),
        "clouds_4k_1step": (clouds, 1, phys, ids + ["POS", "VEL", "PRED_POS", "TOT_CORR_POS", "TEMP", "VAPOR_DENS", "CLOUD_DENS",
                                                    "BUOYANCY", "CLOUD_GEN", "DENSITY", "LAPLACIAN_TEMP", "CORR_TEMP", "PART_ID"]),
    }


def generator_cases():
    """name -> (product generator, reference generator, args): initial-state point sets of the presets (utils/Geometry.cpp)"""
    from realtimeparticles_b200 import _abi
    yz = _abi.PLANE_YZ
    return {
        "rect_yz_64x64_fluids2d_dam": (lambda: _abi.gen_rectangle_grid((64, 64), (0.0, -5.0, -5.0), (0.0, 0.0, 0.0), yz),
                                       lambda: R.generate_2d_grid(0, 2, (64, 64), (0.0, -5.0, -5.0), (0.0, 0.0, 0.0))),
        "circle_yz_32x16_boids2d": (lambda: _abi.gen_circle_grid((32, 16), (0.0, -10 / 6.0, -10 / 6.0), (0.0, 10 / 6.0, 10 / 6.0), yz),
                                    lambda: R.generate_2d_grid(1, 2, (32, 16), (0.0, -10 / 6.0, -10 / 6.0), (0.0, 10 / 6.0, 10 / 6.0))),
        "circle_xy_7x5": (lambda: _abi.gen_circle_grid((7, 5), (-1.5, 0.25, -3.0), (2.5, 1.75, 0.5), _abi.PLANE_XY),
                          lambda: R.generate_2d_grid(1, 0, (7, 5), (-1.5, 0.25, -3.0), (2.5, 1.75, 0.5))),
        "sphere_8x8x8_boids": (lambda: _abi.gen_sphere_grid((8, 8, 8), (-10 / 6.0,) * 3, (10 / 6.0,) * 3),
                               lambda: R.generate_3d_grid(1, (8, 8, 8), (-10 / 6.0,) * 3, (10 / 6.0,) * 3)),
        "random_yz_128x64_clouds2d": (lambda: _abi.gen_random_box(8192, (0.0, -10.0, -5.0), (0.0, 0.0, 5.0), 1),
                                      lambda: R.generate_2d_grid(0, 2, (128, 64), (0.0, -10.0, -5.0), (0.0, 0.0, 5.0), random=True, seed=1)),
    }


def run_case(cls, name):
    make, steps, flags, fields = cases()[name]
    w = make(cls)
    init = {f: w.download(f) for f in ("POS", "VEL")}
    for _ in range(steps):
        w.step(flags)
    return init, {f: w.download(f) for f in fields}


if __name__ == "__main__":
    assert R.available(), "build oracle/_ref first: make -C oracle/ref"
    for name in cases():
        init, out = run_case(R.World, name)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), init_POS=init["POS"], init_VEL=init["VEL"], **out)
        print(name, {k: v.shape for k, v in out.items()})
    np.savez_compressed(os.path.join(OUT, "generators.npz"), **{k: ref() for k, (_, ref) in generator_cases().items()})
    print("generators", list(generator_cases()))
    # boids target trajectory of the reference (Physics::Target): 3D at the default velocity scale, 2D at a faster one
    np.savez_compressed(os.path.join(OUT, "target_trajectory.npz"), dim3_v05=R.target_trajectory(10, 3, 0.5, 400),
                        dim2_v20=R.target_trajectory(10, 2, 2.0, 400))
    print("target_trajectory")
