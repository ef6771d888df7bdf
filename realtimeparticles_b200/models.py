"""Host-side mirror of the reference's physics interface on top of the C ABI.

Same names, argument meaning and error behaviour as
  Physics::Model / ModelParams / CreateModel      physics/Model.hpp:60-216, physics/Model.cpp:10-24
  Physics::CL::Boids / Fluids / Clouds            physics/ocl/{Boids,Fluids,Clouds}.{hpp,cpp}
so the headless harness and the parity tests drive the CUDA backend exactly as the app drives the OpenCL one:
construct with ModelParams, tweak the JSON blob with updateInputJson(), call update() once per frame.
(The C++ drop-in that derives from the unmodified Physics::Model lives in realtimeparticles_b200/cpp/.)
"""
import copy
import enum

import numpy as np

from . import _abi


class ModelType(enum.IntEnum):  # physics/Model.hpp:16-21
    BOIDS = 0
    FLUIDS = 1
    CLOUDS = 2


class Boundary(enum.IntEnum):  # physics/Model.hpp:38-44
    BouncingWall = 0
    CyclicWall = 1


class Dimension(enum.IntEnum):  # utils/Geometry.hpp:11-15
    dim2D = 2
    dim3D = 3


class PhysicsCase(enum.IntEnum):  # utils/Parameters.hpp:17-39
    CASE_INVALID = -1
    BOIDS_BEGIN = 0
    BOIDS_SMALL = 1
    BOIDS_MEDIUM = 2
    BOIDS_LARGE = 3
    BOIDS_XLARGE = 4
    BOIDS_END = 5
    FLUIDS_BEGIN = 6
    FLUIDS_DAM = 7
    FLUIDS_BOMB = 8
    FLUIDS_DROP = 9
    FLUIDS_END = 10
    CLOUDS_BEGIN = 11
    CLOUDS_CUMULUS = 12
    CLOUDS_HOMOGENEOUS = 13
    CLOUDS_END = 14


# utils/Parameters.hpp:61-98
P512, P1K, P4K, P8K, P16K, P32K, P65K, P130K = (1 << 9, 1 << 10, 1 << 12, 1 << 13, 1 << 14, 1 << 15, 1 << 16, 1 << 17)
ALL_NB_PARTICLES = {
    P512: ((32, 16), (8, 8, 8)), P1K: ((32, 32), (16, 8, 8)), P4K: ((64, 64), (16, 16, 16)),
    P8K: ((128, 64), (32, 16, 16)), P16K: ((128, 128), (32, 32, 16)), P32K: ((128, 256), (32, 32, 32)),
    P65K: ((256, 256), (64, 32, 32)), P130K: ((256, 512), (64, 64, 32)),
}


def GetNbParticlesSubdiv2D(nb):
    return ALL_NB_PARTICLES.get(nb, ((0, 0), (0, 0, 0)))[0]


def GetNbParticlesSubdiv3D(nb):
    return ALL_NB_PARTICLES.get(nb, ((0, 0), (0, 0, 0)))[1]


class ModelParams:  # physics/Model.hpp:60-73
    def __init__(self, currNbParticles=0, maxNbParticles=0, boxSize=(0, 0, 0), gridRes=(0, 0, 0), velocity=0.0,
                 particlePosVBO=0, particleColVBO=0, cameraVBO=0, gridVBO=0, dimension=Dimension.dim3D,
                 pCase=PhysicsCase.CASE_INVALID, device=0):
        self.currNbParticles = currNbParticles
        self.maxNbParticles = maxNbParticles
        self.boxSize = tuple(boxSize)
        self.gridRes = tuple(gridRes)
        self.velocity = velocity
        self.particlePosVBO = particlePosVBO
        self.particleColVBO = particleColVBO
        self.cameraVBO = cameraVBO
        self.gridVBO = gridVBO
        self.dimension = dimension
        self.pCase = pCase
        self.device = device  # CUDA ordinal (new; the reference picks the GL-sharing GPU itself)


def _json_diff_empty(a, b):
    return a == b


def _merge_patch(dst, patch):  # RFC 7386, as nlohmann::json::merge_patch
    if not isinstance(patch, dict):
        return copy.deepcopy(patch)
    if not isinstance(dst, dict):
        dst = {}
    for k, v in patch.items():
        if v is None:
            dst.pop(k, None)
        else:
            dst[k] = _merge_patch(dst.get(k), v)
    return dst


class Model:
    """Physics::Model, physics/Model.hpp:81-216 (one CUDA handle per live model)."""

    _TYPE = None
    _MAX_PARTS_IN_CELL = 0

    def __init__(self, params, js=None):
        self.m_maxNbParticles = params.maxNbParticles
        self.m_currNbParticles = params.currNbParticles
        self.m_boxSize = tuple(params.boxSize)
        self.m_gridRes = tuple(params.gridRes)
        self.m_nbCells = params.gridRes[0] * params.gridRes[1] * params.gridRes[2]
        self.m_dimension = Dimension(params.dimension)
        self.m_case = params.pCase
        self.m_boundary = Boundary.BouncingWall
        self.m_init = False
        self.m_pause = False
        self.m_currentDisplayedQuantityName = ""
        self.m_allDisplayableQuantities = {}
        self.m_inputJson = copy.deepcopy(js) if js is not None else {}
        self.m_cameraPos = (32.0, -1.2, 0.0)  # render/Camera.cpp:11; the app feeds it through the camera VBO
        self.m_stepFlags = _abi.STEP_UPDATE
        self._h = _abi.Handle(int(self._TYPE), params.maxNbParticles, params.currNbParticles, params.boxSize,
                              params.gridRes, int(self.m_dimension), self._MAX_PARTS_IN_CELL, params.device)

    # -- typed accessors, physics/Model.hpp:104-176
    def maxNbParticles(self):
        return self.m_maxNbParticles

    def setNbParticles(self, n):
        self.m_currNbParticles = n
        self._h.set_nb_particles(n)

    def nbParticles(self):
        return self.m_currNbParticles

    def setDimension(self, dim):
        self.m_dimension = Dimension(dim)
        self._h.set_dimension(int(self.m_dimension))
        self.reset()

    def dimension(self):
        return self.m_dimension

    def setBoundary(self, boundary):
        self.m_boundary = Boundary(boundary)
        self._h.set_boundary(int(self.m_boundary))

    def boundary(self):
        return self.m_boundary

    def isInit(self):
        return self.m_init

    def pause(self, p):
        self.m_pause = bool(p)

    def onPause(self):
        return self.m_pause

    def targetPos(self):
        return (0.0, 0.0, 0.0)

    def isTargetActivated(self):
        return False

    def isTargetVisible(self):
        return False

    def setCurrentDisplayedQuantity(self, name):  # physics/Model.cpp:26-38
        if name in self.m_allDisplayableQuantities:
            self.m_currentDisplayedQuantityName = name

    def currentDisplayedPhysicalQuantityName(self):
        return self.m_currentDisplayedQuantityName

    def isProfilingEnabled(self):
        return getattr(self, "_profiling", False)

    def enableProfiling(self, enable):
        self._profiling = bool(enable)
        self._h.enable_profiling(enable)

    def isUsingIGPU(self):
        return False

    def getInputJson(self):
        return copy.deepcopy(self.m_inputJson)

    def resetInputJson(self, newJson):
        self.m_inputJson = copy.deepcopy(newJson)

    def updateInputJson(self, newJson):
        if _json_diff_empty(self.m_inputJson, newJson):
            return
        self.m_inputJson = _merge_patch(self.m_inputJson, newJson)
        self.updateModelWithInputJson(self.m_inputJson)

    def setCase(self, c):
        self.m_case = c

    def getCase(self):
        return self.m_case

    # -- OclModel.hpp:41-57
    def updateModelWithInputJson(self, inputJson):
        self.transferJsonInputsToModel(inputJson)
        self.transferKernelInputsToGPU()

    # -- new, harness-facing (the app does this through the GL VBOs)
    def setCameraPos(self, cam):
        self.m_cameraPos = tuple(float(c) for c in cam)

    def setStepFlags(self, flags):
        """Which parts of update() run: physics, render-side kernels, camera sort (default: all, like the app)."""
        self.m_stepFlags = flags

    def upload(self, name, arr, blocking=True):
        """loadBufferFromHost (Context.cpp:355-377); blocking=False: stream-ordered copy from a page-locked array"""
        self._h.upload(name, arr, blocking)

    def download(self, name, out=None, blocking=True):
        """unloadBufferFromDevice (Context.cpp:379-400); blocking=False: `out` (page-locked) is complete after sync()"""
        return self._h.download(name, out, blocking)

    def sync(self):
        self._h.sync()

    def stageTimes(self):
        return self._h.stage_times()

    def handle(self):
        return self._h

    def _flags(self):
        flags = self.m_stepFlags
        if self.m_pause:
            flags &= ~_abi.STEP_PHYSICS
        return flags

    def update(self, replay_graph=True):
        if not self.m_init:
            return
        if self.isProfilingEnabled() or not replay_graph:
            self._h.step(self._flags(), self.m_cameraPos)  # plain launches (with events between the stages when profiling)
        else:
            self._h.step_n(1, self._flags(), self.m_cameraPos)  # the step's cached CUDA graph: one launch, same result

    def updateN(self, n):
        """n consecutive update() calls replayed from one CUDA graph (headless runs)."""
        if not self.m_init:
            return
        self._h.step_n(n, self._flags(), self.m_cameraPos)

    def _load_particles(self, verts, vel=None, col=None):
        M = self.m_maxNbParticles
        pos = np.full((M, 4), np.inf, np.float32)
        pos[:, 3] = 0.0
        k = min(len(verts), M)  # a preset larger than maxNbParticles is cut, like CudaModel::uploadParticles
        pos[:k] = verts[:k]
        self._h.upload("p_pos", pos)
        v = np.zeros((M, 4), np.float32) if vel is None else vel
        self._h.upload("p_vel", v)
        if col is not None:
            c = np.empty((M, 4), np.float32)
            c[:] = col
            self._h.upload("p_col", c)


initBoidsJson = {  # physics/ocl/Boids.cpp:41-73
    "Boids": {
        "Velocity": [0.5, 0.01, 5.0],
        "Target": {"Enable##Target": False, "Show": True, "Radius": [2.0, 1.0, 20.0], "Attract": True},
        "Alignment": {"Enable##Alignment": True, "Scale##Alignment": [1.6, 0.0, 3.0]},
        "Cohesion": {"Enable##Cohesion": True, "Scale##Cohesion": [1.45, 0.0, 3.0]},
        "Separation": {"Enable##Separation": True, "Scale##Separation": [1.6, 0.0, 3.0]},
    }
}

initFluidsJson = {  # physics/ocl/Fluids.cpp:52-74
    "Fluids": {
        "Rest Density": [450.0, 10.0, 1000.0],
        "Relax CFM": [600.0, 100.0, 1000.0],
        "Time Step": [0.010, 0.0001, 0.020],
        "Nb Jacobi Iterations": [2, 1, 6],
        "Artificial Pressure": {"Enable##Pressure": True, "Coefficient##Pressure": [0.001, 0.0, 0.001],
                                "Radius": [0.006, 0.001, 0.015], "Exp": [4, 1, 6]},
        "Vorticity Confinement": {"Enable##Vorticity": True, "Coefficient##Vorticity": [0.0004, 0.0, 0.001],
                                  "xSPH Viscosity Coefficient": [0.0001, 0.0, 0.001]},
    }
}

initCloudsJson = {  # physics/ocl/Clouds.cpp:67-102
    "Fluids": copy.deepcopy(initFluidsJson["Fluids"]),
    "Clouds": {
        "Enable Temperature Smoothing": True,
        "Ground Heat Coefficient": [10.0, 0.0, 1000.0],
        "Buoyancy Heat Coefficient": [0.10, 0.0, 5.0],
        "Gravity Coefficient": [0.0005, 0.0, 0.1],
        "Adiabatic Lapse Rate": [5.0, 0.0, 20.0],
        "Phase Transition Rate": [0.3485, 0.0, 20.0],
        "Latent Heat Coefficient": [0.07, 0.0, 0.100],
        "Wind Coefficient": [1.0, 0.0, 1.0],
    },
}


def _fluid_inputs_from_json(fluidsJson, dim, k):
    k.restDensity = float(fluidsJson["Rest Density"][0])
    k.relaxCFM = float(fluidsJson["Relax CFM"][0])
    k.timeStep = float(fluidsJson["Time Step"][0])
    k.dim = 2 if dim == Dimension.dim2D else 3
    ap = fluidsJson["Artificial Pressure"]
    k.isArtPressureEnabled = 1 if ap["Enable##Pressure"] is True else 0
    k.artPressureCoeff = float(ap["Coefficient##Pressure"][0])
    k.artPressureRadius = float(ap["Radius"][0])
    k.artPressureExp = int(ap["Exp"][0])
    vc = fluidsJson["Vorticity Confinement"]
    k.isVorticityConfEnabled = 1 if vc["Enable##Vorticity"] is True else 0
    k.vorticityConfCoeff = float(vc["Coefficient##Vorticity"][0])
    k.xsphViscosityCoeff = float(vc["xSPH Viscosity Coefficient"][0])


class Boids(Model):
    """Physics::CL::Boids, physics/ocl/Boids.{hpp,cpp}. The Perlin-noise target trajectory (physics/utils/Target.cpp) is
    evaluated on the host once per update(), like the reference (Boids.cpp:351-358): rtp_target_update()."""

    _TYPE = ModelType.BOIDS
    _MAX_PARTS_IN_CELL = 3000  # Boids.cpp:78

    def __init__(self, params):
        super().__init__(params, initBoidsJson)
        self.m_rules = _abi.BoidsParams(0.5, 1.6, 1.6, 1.45)
        self.m_targetInputs = _abi.TargetParams(2.0, 1)
        self.m_targetActive = False
        self.m_targetVisible = False
        self.m_targetPos = (0.0, 0.0, 0.0)
        self.m_target = _abi.Target(int(params.boxSize[0]))
        self.m_init = True
        self.reset()

    def update(self):  # Boids.cpp:323-384
        if not self.m_init:
            return
        if not self.m_pause and self.m_targetActive:
            self.m_targetPos = self.m_target.update(3 if self.m_dimension == Dimension.dim3D else 2, self.m_rules.velocityScale)
            self.transferKernelInputsToGPU()
            super().update(replay_graph=False)  # (the target moves every frame: new kernel parameters, nothing to replay)
            return
        super().update()

    def transferJsonInputsToModel(self, inputJson):  # Boids.cpp:177-210
        if not self.m_init:
            return
        try:
            b = inputJson["Boids"]
            self.m_rules.velocityScale = float(b["Velocity"][0])
            self.m_rules.alignmentScale = float(b["Alignment"]["Scale##Alignment"][0]) if b["Alignment"]["Enable##Alignment"] else 0.0
            self.m_rules.separationScale = float(b["Separation"]["Scale##Separation"][0]) if b["Separation"]["Enable##Separation"] else 0.0
            self.m_rules.cohesionScale = float(b["Cohesion"]["Scale##Cohesion"][0]) if b["Cohesion"]["Enable##Cohesion"] else 0.0
            self.m_targetActive = bool(b["Target"]["Enable##Target"])
            self.m_targetVisible = bool(b["Target"]["Show"])
            self.m_targetInputs.targetRadiusEffect = float(b["Target"]["Radius"][0])
            self.m_targetInputs.targetSignEffect = int(bool(b["Target"]["Attract"]))  # (int)bool, Boids.cpp:201: "repel" is 0
        except (KeyError, TypeError, IndexError):
            raise RuntimeError("Wrong Json parsing")

    def transferKernelInputsToGPU(self):  # Boids.cpp:212-226
        self._h.set_boids_params(self.m_rules, self.m_targetInputs, tuple(self.m_targetPos) + (0.0,), self.m_targetActive)

    def isTargetActivated(self):
        return self.m_targetActive

    def isTargetVisible(self):
        return self.m_targetVisible

    def targetPos(self):
        return self.m_targetPos

    def setTargetPos(self, pos):
        self.m_targetPos = tuple(float(p) for p in pos)
        self.transferKernelInputsToGPU()

    def reset(self):  # Boids.cpp:228-275
        if not self.m_init:
            return
        self.resetInputJson(initBoidsJson)
        nb = {PhysicsCase.BOIDS_SMALL: P512, PhysicsCase.BOIDS_MEDIUM: P16K, PhysicsCase.BOIDS_LARGE: P65K,
              PhysicsCase.BOIDS_XLARGE: P130K}.get(self.m_case)
        if nb is not None:
            self.setNbParticles(nb)
        self.updateModelWithInputJson(self.m_inputJson)
        self.initBoidsParticles()
        self._h.upload("p_col", np.tile(np.array([1.0, 0.02, 0.02, 0.5], np.float32), (self.m_maxNbParticles, 1)))
        self._h.reset_ids()

    def initBoidsParticles(self):  # Boids.cpp:277-321
        if self.m_currNbParticles > self.m_maxNbParticles:
            return
        bx, by, bz = [float(b) for b in self.m_boxSize]
        if self.m_dimension == Dimension.dim2D:  # circle in the YZ plane, Boids.cpp:291-298
            sub = GetNbParticlesSubdiv2D(self.m_currNbParticles)
            verts = _abi.gen_circle_grid(sub, (0.0, by / -6.0, bz / -6.0), (0.0, by / 6.0, bz / 6.0), _abi.PLANE_YZ)
        else:
            sub = GetNbParticlesSubdiv3D(self.m_currNbParticles)
            verts = _abi.gen_sphere_grid(sub, (bx / -6.0, by / -6.0, bz / -6.0), (bx / 6.0, by / 6.0, bz / 6.0))
        M = self.m_maxNbParticles
        pos = np.full((M, 4), np.inf, np.float32)
        pos[:, 3] = 0.0
        k = min(len(verts), M)  # a preset larger than maxNbParticles is cut, like CudaModel::uploadParticles
        pos[:k] = verts[:k]
        self._h.upload("p_pos", pos)
        self._h.upload("p_vel", pos)  # "Using same buffer to initialize vel", Boids.cpp:316-318


class Fluids(Model):
    """Physics::CL::Fluids, physics/ocl/Fluids.{hpp,cpp}."""

    _TYPE = ModelType.FLUIDS
    _MAX_PARTS_IN_CELL = 100  # Fluids.cpp:79

    def __init__(self, params):
        super().__init__(params, initFluidsJson)
        self.m_kernelInputs = _abi.FluidParams(450.0, 600.0, 0.010, 3, 1, 0.006, 0.001, 4, 1, 0.0004, 0.0001)
        self.m_nbJacobiIters = 2
        self.m_init = True
        self.reset()

    def transferJsonInputsToModel(self, inputJson):  # Fluids.cpp:218-251
        if not self.m_init or not inputJson:
            return
        try:
            f = inputJson["Fluids"]
            self.m_nbJacobiIters = int(f["Nb Jacobi Iterations"][0])
            _fluid_inputs_from_json(f, self.m_dimension, self.m_kernelInputs)
        except (KeyError, TypeError, IndexError):
            raise RuntimeError("Wrong Json parsing")

    def transferKernelInputsToGPU(self):  # Fluids.cpp:253-271
        if not self.m_init:
            return
        self._h.set_fluid_params(self.m_kernelInputs, self.m_nbJacobiIters)

    def reset(self):  # Fluids.cpp:196-216
        if not self.m_init:
            return
        self.resetInputJson(initFluidsJson)
        self.updateModelWithInputJson(self.m_inputJson)
        self.initFluidsParticles()
        self._h.reset_ids()

    def initFluidsParticles(self):  # Fluids.cpp:273-398
        bx, by, bz = [float(b) for b in self.m_boxSize]
        if self.m_dimension == Dimension.dim2D:  # rectangles in the YZ plane, Fluids.cpp:287-335
            if self.m_case == PhysicsCase.FLUIDS_DAM:
                nb, start, end = P4K, (0.0, by / -2.0, bz / -2.0), (0.0, 0.0, 0.0)
            elif self.m_case == PhysicsCase.FLUIDS_BOMB:
                nb, start, end = P4K, (0.0, by / -6.0, bz / -6.0), (0.0, by / 6.0, bz / 6.0)
            elif self.m_case == PhysicsCase.FLUIDS_DROP:
                nb, start, end = P512, (0.0, 2.0 * by / 10.0, bz / -10.0), (0.0, 4.0 * by / 10.0, bz / 10.0)
            else:
                return
            verts = _abi.gen_rectangle_grid(GetNbParticlesSubdiv2D(nb), start, end, _abi.PLANE_YZ)
            if self.m_case == PhysicsCase.FLUIDS_DROP:  # + the pool it falls into
                nb += P4K
                pool = _abi.gen_rectangle_grid((64, 128), (0.0, by / -2.0, bz / -2.0), (0.0, 0.0, bz / 2.0), _abi.PLANE_YZ)
                verts = np.concatenate([verts, pool])
            self.setNbParticles(nb)
            self._load_particles(verts, col=np.array([0.0, 0.1, 1.0, 0.0], np.float32))
            return
        if self.m_case == PhysicsCase.FLUIDS_DAM:
            nb, box = P130K, True
            start, end = (bx / -2.0, by / -2.0, bz / -2.0), (bx / 2.0, 0.0, 0.0)
        elif self.m_case == PhysicsCase.FLUIDS_BOMB:
            nb, box = P65K, False
            start, end = (bx / -6.0, by / -6.0, bz / -6.0), (bx / 6.0, by / 6.0, bz / 6.0)
        elif self.m_case == PhysicsCase.FLUIDS_DROP:
            nb, box = P4K, True
            start, end = (bx / -10.0, 2.0 * by / 10.0, bz / -10.0), (bx / 10.0, 4.0 * by / 10.0, bz / 10.0)
        else:
            return  # "Unkown case type": keep what the caller uploaded
        gen = _abi.gen_box_grid if box else _abi.gen_sphere_grid
        verts = gen(GetNbParticlesSubdiv3D(nb), start, end)
        if self.m_case == PhysicsCase.FLUIDS_DROP:
            nb += P65K
            floor = _abi.gen_box_grid((64, 16, 64), (bx / -2.0, by / -2.0, bz / -2.0), (bx / 2.0, by / -2.55, bz / 2.0))
            verts = np.concatenate([verts, floor])
        self.setNbParticles(nb)
        self._load_particles(verts, col=np.array([0.0, 0.1, 1.0, 0.0], np.float32))


class Clouds(Model):
    """Physics::CL::Clouds, physics/ocl/Clouds.{hpp,cpp}."""

    _TYPE = ModelType.CLOUDS
    _MAX_PARTS_IN_CELL = 100  # Clouds.cpp:107

    def __init__(self, params):
        super().__init__(params, initCloudsJson)
        self.m_fluidKernelInputs = _abi.FluidParams(450.0, 600.0, 0.010, 3, 1, 0.006, 0.001, 4, 1, 0.0004, 0.0001)
        self.m_cloudKernelInputs = _abi.CloudParams(3, 0.01, 400.0, 10.0, 0.10, 0.0005, 5.0, 0.3485, 0.07, 1, 600.0, 0.75, 1.0)
        self.m_nbJacobiIters = 1
        M = self.m_maxNbParticles
        # Clouds.cpp:202-213: name -> (buffer, static range, user range)
        self.m_allDisplayableQuantities = {
            "Particle ID": ["p_partID", (0.0, float(M - 1)), (0.0, float(32000 - 1))],
            "Vapor Density": ["p_vaporDens", (0.0, 100.0), (0.001, 100.0)],
            "Cloud Density": ["p_cloudDens", (0.0, 100.0), (1.0, 15.0)],
            "Net Force": ["p_buoyancy", (-10.0, 10.0), (-1.0, 1.0)],
            "Temperature": ["p_temp", (0.0, 500.0), (223.0, 293.0)],
        }
        self.m_currentDisplayedQuantityName = "Cloud Density"
        self.m_init = True
        self.reset()

    def transferJsonInputsToModel(self, inputJson):  # Clouds.cpp:279-329
        if not self.m_init:
            return
        try:
            f = inputJson["Fluids"]
            self.m_nbJacobiIters = int(f["Nb Jacobi Iterations"][0])
            _fluid_inputs_from_json(f, self.m_dimension, self.m_fluidKernelInputs)
            c = inputJson["Clouds"]
            k = self.m_cloudKernelInputs
            k.restDensity = float(f["Rest Density"][0])
            k.timeStep = float(f["Time Step"][0])
            k.dim = 2 if self.m_dimension == Dimension.dim2D else 3
            k.relaxCFM = float(f["Relax CFM"][0])
            k.isTempSmoothingEnabled = 1 if c["Enable Temperature Smoothing"] else 0
            k.groundHeatCoeff = float(c["Ground Heat Coefficient"][0])
            k.buoyancyCoeff = float(c["Buoyancy Heat Coefficient"][0])
            k.gravCoeff = float(c["Gravity Coefficient"][0])
            k.adiabaticLapseRate = float(c["Adiabatic Lapse Rate"][0])
            k.phaseTransitionRate = float(c["Phase Transition Rate"][0])
            k.latentHeatCoeff = float(c["Latent Heat Coefficient"][0])
            k.windCoeff = float(c["Wind Coefficient"][0])
        except (KeyError, TypeError, IndexError):
            raise RuntimeError("Wrong Json parsing")

    def transferKernelInputsToGPU(self):  # Clouds.cpp:331-382
        if not self.m_init:
            return
        self._h.set_fluid_params(self.m_fluidKernelInputs, self.m_nbJacobiIters)
        self._h.set_cloud_params(self.m_cloudKernelInputs)

    def setCurrentDisplayedQuantity(self, name):
        super().setCurrentDisplayedQuantity(name)
        self._push_displayed_quantity()

    def _push_displayed_quantity(self):
        q = self.m_allDisplayableQuantities.get(self.m_currentDisplayedQuantityName)
        if q:
            self._h.set_displayed_quantity(q[0], q[2][0], q[2][1])

    def reset(self):  # Clouds.cpp:384-399
        if not self.m_init:
            return
        self.resetInputJson(initCloudsJson)
        self.updateModelWithInputJson(self.m_inputJson)
        self.initCloudsParticles()
        self._push_displayed_quantity()
        self._h.reset_ids()

    def initCloudsParticles(self):  # Clouds.cpp:401-501
        bx, by, bz = [float(b) for b in self.m_boxSize]
        if self.m_dimension == Dimension.dim2D:  # random fill of a YZ rectangle (x extent 0), Clouds.cpp:414-444
            if self.m_case == PhysicsCase.CLOUDS_CUMULUS:
                nb, start, end = P8K, (0.0, by / -2.0, bz / -2.0), (0.0, 0.0, bz / 2.0)
            elif self.m_case == PhysicsCase.CLOUDS_HOMOGENEOUS:
                nb, start, end = P8K, (0.0, by / -2.0, bz / -2.0), (0.0, by / 2.0, bz / 2.0)
            else:
                return
        elif self.m_case == PhysicsCase.CLOUDS_CUMULUS:
            nb, start, end = P65K, (bx / -2.0, by / -2.0, bz / -2.0), (bx / 2.0, by / -4.0, bz / 2.0)
        elif self.m_case == PhysicsCase.CLOUDS_HOMOGENEOUS:
            nb, start, end = P65K, (bx / -2.0, by / -2.0, bz / -2.0), (bx / 2.0, by / 2.0, bz / 2.0)
        else:
            return
        verts = _abi.gen_random_box(nb, start, end, seed=-1)  # unseeded rand(): continues the process' sequence
        self.setNbParticles(nb)
        self.loadCloudsState(verts)

    def loadCloudsState(self, verts):
        """Clouds.cpp:474-499: positions, zero velocity, zero cloud density, particle ids, then the two init kernels."""
        M = self.m_maxNbParticles
        self._load_particles(verts, col=np.array([0.0, 0.1, 1.0, 0.0], np.float32))
        self._h.upload("p_cloudDens", np.zeros(M, np.float32))
        self._h.upload("p_partID", np.arange(M, dtype=np.float32))
        self._h.init_clouds_fields()


def CreateModel(type, params):  # physics/Model.cpp:10-24
    t = int(type)
    if t == ModelType.BOIDS:
        return Boids(params)
    if t == ModelType.FLUIDS:
        return Fluids(params)
    if t == ModelType.CLOUDS:
        return Clouds(params)
    return None
