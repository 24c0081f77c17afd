"""rawphys_b200 -- ctypes binding of librawphys_b200.so (include/rawphys_b200.h), the B200-native XPBD frame step that
stands in for raw-physics' pbd_simulate / pbd_simulate_with_constraints (src/physics/pbd.h:93-94).

This module is plumbing for tests and the benchmark; the product is the shared library. There is NO fallback: if the
library has not been built (`python raw-physics_b200/build.py`) importing this module raises, and every batch call
needs a CUDA device.

The directory name contains a hyphen, so load it through `__graft_entry__.load_package()` (module name rawphys_b200).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RAWPHYS_B200_LIB") or os.path.join(HERE, "librawphys_b200.so")  # override: tuning variants
STATE_STRIDE = 21
PARAM_STRIDE = 25

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)

# every symbol include/rawphys_b200.h declares (tests check the library exports each one)
EXPORTS = [
    "rp_last_error", "rp_device_count", "rp_scene_create", "rp_scene_destroy", "rp_scene_collider_hull", "rp_scene_collider_sphere",
    "rp_scene_add_body", "rp_scene_collider_hull_topology", "rp_scene_add_body_params", "rp_scene_add_positional_constraint", "rp_scene_add_mutual_orientation_constraint",
    "rp_scene_add_hinge_joint_constraint", "rp_scene_add_spherical_joint_constraint", "rp_scene_num_bodies", "rp_scene_get_params",
    "rp_scene_hull_sizes", "rp_scene_hull_dump", "rp_batch_cfg_default", "rp_batch_create", "rp_batch_destroy", "rp_batch_num_worlds",
    "rp_batch_num_bodies", "rp_batch_clear_forces", "rp_batch_add_force", "rp_batch_add_gravity", "rp_batch_step", "rp_batch_sync",
    "rp_batch_run", "rp_batch_upload_state", "rp_batch_download_state", "rp_batch_broadcast_state", "rp_batch_step_host",
    "rp_batch_get_status", "rp_batch_clear_status", "rp_batch_get_counters", "rp_batch_step_logged", "rp_batch_broad_pairs", "rp_batch_profile",
    "rp_measure_fp64_peak", "rp_scene_initial_state", "rp_scene_body_desc", "rp_scene_collider_soup_size", "rp_scene_collider_soup",
    "rp_scene_num_joints", "rp_scene_joint_desc", "rp_batch_graph_kernels", "rp_batch_pair_levels", "rp_batch_create_from", "rp_example_count", "rp_example_name", "rp_example_error", "rp_example_create",
    "rp_example_create_on", "rp_scene_set_hull_device", "rp_scene_hull_build_stats",
]
KERNEL_FAMILIES = ["broadphase", "islands", "schedule", "integrate", "cull", "gjk", "manifold", "solve_pos", "derive", "solve_vel", "epa"]


class BatchCfg(C.Structure):
    _fields_ = [("max_pairs_per_world", C.c_uint32), ("max_contacts_per_world", C.c_uint32), ("disable_cull", C.c_uint32),
                ("solve_order", C.c_uint32), ("sweep_block_worlds", C.c_uint32), ("large_scene", C.c_uint32), ("disable_islands", C.c_uint32), ("sweep_form", C.c_uint32),
                ("linear_sleeping_threshold", C.c_double), ("angular_sleeping_threshold", C.c_double), ("deactivation_time", C.c_double)]


class ExampleInfo(C.Structure):
    _fields_ = [("substeps", C.c_uint32), ("pos_iters", C.c_uint32), ("collisions", C.c_int32), ("reserved0", C.c_int32), ("gravity", C.c_double)]


class RawPhysError(RuntimeError):
    pass


class RawPhysCapacityError(RawPhysError):
    """RP_ERR_CAPACITY: a fixed-capacity device buffer ran out in some world (pairs / contacts were dropped there)"""


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RawPhysError("librawphys_b200.so is not built (run `python raw-physics_b200/build.py`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.rp_last_error.restype = C.c_char_p
    L.rp_scene_create.restype = C.c_void_p
    L.rp_scene_destroy.argtypes = [C.c_void_p]
    L.rp_scene_collider_hull.argtypes = [C.c_void_p, _dp, C.c_uint32, _u32p, C.c_uint32]
    L.rp_scene_collider_sphere.argtypes = [C.c_void_p, C.c_float]
    L.rp_scene_add_body.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double]
    L.rp_scene_collider_hull_topology.argtypes = [C.c_void_p, _dp, C.c_uint32, _dp, C.c_uint32] + [_u32p] * 8
    L.rp_scene_add_body_params.argtypes = [C.c_void_p, _dp, _dp, C.c_double, _dp, _dp, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double]
    L.rp_scene_add_positional_constraint.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_double, _dp]
    L.rp_scene_add_mutual_orientation_constraint.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    L.rp_scene_add_hinge_joint_constraint.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_int,
                                                       C.c_int, C.c_int, C.c_double, C.c_double]
    L.rp_scene_add_spherical_joint_constraint.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                                           C.c_double, C.c_double, C.c_double, C.c_double]
    L.rp_scene_num_bodies.argtypes = [C.c_void_p]
    L.rp_scene_get_params.argtypes = [C.c_void_p, _dp]
    L.rp_scene_hull_sizes.argtypes = [C.c_void_p, C.c_int, C.c_int, _i32p]
    L.rp_scene_hull_dump.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp] + [_u32p] * 8
    L.rp_batch_cfg_default.argtypes = [C.POINTER(BatchCfg)]
    L.rp_batch_create.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.POINTER(BatchCfg), C.POINTER(C.c_void_p)]
    L.rp_batch_destroy.argtypes = [C.c_void_p]
    L.rp_batch_create_from.argtypes = [C.c_void_p, C.c_void_p, _i32p, C.POINTER(BatchCfg), C.POINTER(C.c_void_p)]
    L.rp_batch_num_worlds.argtypes = [C.c_void_p]
    L.rp_batch_num_worlds.restype = C.c_uint32
    L.rp_batch_num_bodies.argtypes = [C.c_void_p]
    L.rp_batch_num_bodies.restype = C.c_uint32
    L.rp_batch_clear_forces.argtypes = [C.c_void_p]
    L.rp_batch_add_force.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    L.rp_batch_add_gravity.argtypes = [C.c_void_p, C.c_double]
    L.rp_batch_step.argtypes = [C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_int]
    L.rp_batch_sync.argtypes = [C.c_void_p]
    L.rp_batch_graph_kernels.argtypes = [C.c_void_p]
    L.rp_batch_run.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_float)]
    L.rp_batch_upload_state.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.rp_batch_download_state.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.rp_batch_broadcast_state.argtypes = [C.c_void_p, C.c_void_p]
    L.rp_batch_step_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_int]
    L.rp_batch_get_status.argtypes = [C.c_void_p, _i32p]
    L.rp_batch_clear_status.argtypes = [C.c_void_p]
    L.rp_batch_get_counters.argtypes = [C.c_void_p, _u64p]
    L.rp_batch_step_logged.argtypes = [C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, _u32p, C.c_uint32, _dp,
                                       C.c_uint32, _u32p, _u32p]
    L.rp_batch_broad_pairs.argtypes = [C.c_void_p, C.c_uint32, _u32p, C.c_uint32, _u32p]
    L.rp_batch_pair_levels.argtypes = [C.c_void_p, C.c_uint32, _u32p, _i32p, C.c_uint32, _u32p]
    L.rp_batch_profile.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_float)]
    L.rp_measure_fp64_peak.argtypes = [C.c_int, _dp]
    L.rp_scene_initial_state.argtypes = [C.c_void_p, _dp]
    L.rp_scene_body_desc.argtypes = [C.c_void_p, C.c_int, _dp]
    L.rp_scene_collider_soup_size.argtypes = [C.c_void_p, C.c_int, C.c_int, _u32p, _u32p, C.POINTER(C.c_float)]
    L.rp_scene_collider_soup.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _u32p]
    L.rp_scene_num_joints.argtypes = [C.c_void_p]
    L.rp_scene_joint_desc.argtypes = [C.c_void_p, C.c_int, _i32p, _dp]
    L.rp_example_name.restype = C.c_char_p
    L.rp_example_name.argtypes = [C.c_int]
    L.rp_example_error.restype = C.c_char_p
    L.rp_example_create.restype = C.c_void_p
    L.rp_example_create.argtypes = [C.c_char_p, _dp, C.c_uint32, C.c_int, C.c_char_p, C.POINTER(ExampleInfo)]
    L.rp_example_create_on.restype = C.c_void_p
    L.rp_example_create_on.argtypes = [C.c_char_p, _dp, C.c_uint32, C.c_int, C.c_char_p, C.POINTER(ExampleInfo), C.c_int]
    L.rp_scene_set_hull_device.argtypes = [C.c_void_p, C.c_int]
    L.rp_scene_hull_build_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        cls = RawPhysCapacityError if rc == 3 else RawPhysError
        raise cls("%s failed (code %d): %s" % (what, rc, lib().rp_last_error().decode()))


def _d(a):
    return a.ctypes.data_as(_dp)


def _u(a):
    return a.ctypes.data_as(_u32p)


def _vec(x):
    return np.ascontiguousarray(x, dtype=np.float64)


class ColliderDescPy:
    def __init__(self, kind, vertices, indices, radius):
        self.kind, self.vertices, self.indices, self.radius = kind, vertices, indices, radius


class BodyDescPy:
    def __init__(self, position, rotation, mass, fixed, colliders, mu_s, mu_d, restitution):
        self.position, self.rotation, self.mass, self.fixed, self.colliders = position, rotation, mass, fixed, colliders
        self.mu_s, self.mu_d, self.restitution = mu_s, mu_d, restitution


class SceneDesc:
    """A scene description with the attribute names of tests/scenes.py's Scene (see Scene.describe)."""

    def __init__(self, name, substeps=20, iters=1, collisions=True, gravity=10.0):
        self.name, self.substeps, self.iters, self.collisions, self.gravity = name, substeps, iters, collisions, gravity
        self.bodies, self.constraints, self.forces, self.initial_state = [], [], [], None


def example_names():
    L = lib()
    return [L.rp_example_name(i).decode() for i in range(L.rp_example_count())]


def example(name, params=(), perturb=False, mesh_dir=None, hull_device=-1):
    """A built-in scene (rp_example_create): -> (Scene, SceneDesc). The description carries the example's own step settings
    (what its update() passes to pbd_simulate) and can be loaded into the oracle. hull_device >= 0: hull topology is built on
    that GPU (rp_scene_set_hull_device) instead of the host."""
    L = lib()
    p = np.ascontiguousarray(params, dtype=np.float64)
    info = ExampleInfo()
    mesh_dir = mesh_dir or os.path.join(HERE, "assets", "meshes")  # (explicit: a tuning variant of the library lives elsewhere)
    h = L.rp_example_create_on(name.encode(), _d(p) if p.size else None, int(p.size), int(perturb), mesh_dir.encode(), C.byref(info),
                               int(hull_device))
    if not h:
        raise RawPhysError("rp_example_create(%r): %s" % (name, L.rp_example_error().decode()))
    sc = Scene(handle=h)
    return sc, sc.describe(name, int(info.substeps), int(info.pos_iters), bool(info.collisions), float(info.gravity))


class Scene:
    """Scene template (rp_scene). `desc` is a description object with .bodies (position, rotation xyzw, mass, fixed,
    colliders[kind, vertices, indices, radius], mu_s, mu_d, restitution) and .constraints (dicts), as tests/scenes.py builds."""

    def __init__(self, desc=None, handle=None, hull_device=-1):
        self.L = lib()
        self.h = C.c_void_p(handle if handle is not None else self.L.rp_scene_create())
        if hull_device >= 0 and self.L.rp_scene_set_hull_device(self.h, int(hull_device)) != 0:
            raise RawPhysError("rp_scene_set_hull_device: %s" % self.L.rp_last_error().decode())
        if desc is not None:
            self.load(desc)

    def hull_build_stats(self):
        """(hulls built, milliseconds spent building them)"""
        n, ms = C.c_int(0), C.c_double(0.0)
        self.L.rp_scene_hull_build_stats(self.h, C.byref(n), C.byref(ms))
        return n.value, ms.value

    def __del__(self):
        if getattr(self, "h", None):
            self.L.rp_scene_destroy(self.h)
            self.h = None

    def load(self, desc):
        L = self.L
        for b in desc.bodies:
            for col in b.colliders:
                if col.kind == "sphere":
                    if L.rp_scene_collider_sphere(self.h, C.c_float(col.radius)) < 0:
                        raise RawPhysError("rp_scene_collider_sphere failed")
                else:
                    v = np.ascontiguousarray(col.vertices, dtype=np.float64)
                    idx = np.ascontiguousarray(col.indices, dtype=np.uint32)
                    if L.rp_scene_collider_hull(self.h, _d(v), v.shape[0], _u(idx), idx.shape[0]) < 0:
                        raise RawPhysError("rp_scene_collider_hull failed")
            if L.rp_scene_add_body(self.h, _d(_vec(b.position)), _d(_vec(b.rotation)), b.mass, int(b.fixed), b.mu_s, b.mu_d,
                                   b.restitution) < 0:
                raise RawPhysError("rp_scene_add_body failed")
        for c in desc.constraints:
            r1 = _vec(c.get("r1", (0, 0, 0)))
            r2 = _vec(c.get("r2", (0, 0, 0)))
            if c["type"] == "positional":
                rc = L.rp_scene_add_positional_constraint(self.h, c["e1"], c["e2"], _d(r1), _d(r2), c["compliance"], _d(_vec(c["distance"])))
            elif c["type"] == "mutual_orientation":
                rc = L.rp_scene_add_mutual_orientation_constraint(self.h, c["e1"], c["e2"], c["compliance"])
            elif c["type"] == "hinge":
                rc = L.rp_scene_add_hinge_joint_constraint(self.h, c["e1"], c["e2"], _d(r1), _d(r2), c["compliance"], c["e1_aligned"],
                                                           c["e2_aligned"], int(c["limited"]), c.get("e1_limit", 0), c.get("e2_limit", 0),
                                                           c.get("lower", 0.0), c.get("upper", 0.0))
            elif c["type"] == "spherical":
                rc = L.rp_scene_add_spherical_joint_constraint(self.h, c["e1"], c["e2"], _d(r1), _d(r2), c["e1_swing"], c["e2_swing"],
                                                               c["e1_twist"], c["e2_twist"], c["swing_lower"], c["swing_upper"],
                                                               c["twist_lower"], c["twist_upper"])
            else:
                raise ValueError(c["type"])
            if rc < 0:
                raise RawPhysError("adding constraint %r failed" % (c,))
        return self

    @property
    def n(self):
        return int(self.L.rp_scene_num_bodies(self.h))

    def initial_state(self):
        """[n, STATE_STRIDE] records the worlds of a batch start from"""
        out = np.zeros((self.n, STATE_STRIDE))
        _check(self.L.rp_scene_initial_state(self.h, _d(out)), "rp_scene_initial_state")
        return out

    def describe(self, name="scene", substeps=20, iters=1, collisions=True, gravity=10.0):
        """The scene as a description object with the attributes of tests/scenes.py's Scene (bodies with their collider soups,
        constraint dicts, step settings, initial state): what refdrv.RefWorld.load() and Scene.load() take."""
        L = self.L
        d = SceneDesc(name=name, substeps=substeps, iters=iters, collisions=collisions, gravity=gravity)
        for b in range(self.n):
            o = np.zeros(16)
            _check(L.rp_scene_body_desc(self.h, b, _d(o)), "rp_scene_body_desc")
            cols = []
            for c in range(int(o[12])):
                nv, ni, rad = C.c_uint32(), C.c_uint32(), C.c_float()
                _check(L.rp_scene_collider_soup_size(self.h, b, c, C.byref(nv), C.byref(ni), C.byref(rad)), "rp_scene_collider_soup_size")
                if nv.value == 0:
                    cols.append(ColliderDescPy("sphere", None, None, float(rad.value)))
                else:
                    v = np.zeros((nv.value, 3))
                    idx = np.zeros(ni.value, dtype=np.uint32)
                    _check(L.rp_scene_collider_soup(self.h, b, c, _d(v), _u(idx)), "rp_scene_collider_soup")
                    cols.append(ColliderDescPy("hull", v, idx, 0.0))
            d.bodies.append(BodyDescPy(tuple(o[0:3]), tuple(o[3:7]), float(o[7]), bool(o[8]), cols, float(o[9]), float(o[10]), float(o[11])))
        names = ["e1_aligned", "e2_aligned", "e1_limit", "e2_limit"]
        for j in range(int(L.rp_scene_num_joints(self.h))):
            iv = np.zeros(8, dtype=np.int32)
            dv = np.zeros(14)
            _check(L.rp_scene_joint_desc(self.h, j, iv.ctypes.data_as(_i32p), _d(dv)), "rp_scene_joint_desc")
            t, e1, e2 = int(iv[0]), int(iv[1]), int(iv[2])
            r1, r2 = tuple(dv[0:3]), tuple(dv[3:6])
            if t == 0:
                c = dict(type="positional", e1=e1, e2=e2, r1=r1, r2=r2, compliance=float(dv[9]), distance=tuple(dv[6:9]))
            elif t == 2:
                c = dict(type="mutual_orientation", e1=e1, e2=e2, compliance=float(dv[9]))
            elif t == 3:
                c = dict(type="hinge", e1=e1, e2=e2, r1=r1, r2=r2, compliance=float(dv[9]), limited=bool(iv[3]), lower=float(dv[10]), upper=float(dv[11]))
                c.update({k: int(iv[4 + i]) for i, k in enumerate(names)})
            else:
                c = dict(type="spherical", e1=e1, e2=e2, r1=r1, r2=r2, e1_swing=int(iv[4]), e2_swing=int(iv[5]), e1_twist=int(iv[6]), e2_twist=int(iv[7]),
                         swing_lower=float(dv[10]), swing_upper=float(dv[11]), twist_lower=float(dv[12]), twist_upper=float(dv[13]))
            d.constraints.append(c)
        st = self.initial_state()
        if np.any(st[:, 7:13] != 0.0):
            d.initial_state = st[:, :15].copy()
        return d

    def params(self):
        out = np.zeros((self.n, PARAM_STRIDE))
        _check(self.L.rp_scene_get_params(self.h, _d(out)), "rp_scene_get_params")
        return out

    def hull(self, body, collider=0):
        sizes = np.zeros(6, dtype=np.int32)
        _check(self.L.rp_scene_hull_sizes(self.h, body, collider, sizes.ctypes.data_as(_i32p)), "rp_scene_hull_sizes")
        if sizes[0] < 0:
            return None
        V, F, fe, v2f, v2n, f2n = [int(x) for x in sizes]
        h = dict(verts=np.zeros((V, 3)), normals=np.zeros((F, 3)),
                 face_ptr=np.zeros(F + 1, np.uint32), face_idx=np.zeros(fe, np.uint32),
                 v2f_ptr=np.zeros(V + 1, np.uint32), v2f_idx=np.zeros(v2f, np.uint32),
                 v2n_ptr=np.zeros(V + 1, np.uint32), v2n_idx=np.zeros(v2n, np.uint32),
                 f2n_ptr=np.zeros(F + 1, np.uint32), f2n_idx=np.zeros(f2n, np.uint32))
        _check(self.L.rp_scene_hull_dump(self.h, body, collider, _d(h["verts"]), _d(h["normals"]), _u(h["face_ptr"]), _u(h["face_idx"]),
                                         _u(h["v2f_ptr"]), _u(h["v2f_idx"]), _u(h["v2n_ptr"]), _u(h["v2n_idx"]), _u(h["f2n_ptr"]),
                                         _u(h["f2n_idx"])), "rp_scene_hull_dump")
        return h


class Batch:
    """n_worlds instances of a scene on one GPU (rp_batch)."""

    def __init__(self, scene, n_worlds=1, device=0, max_pairs=0, max_contacts=0, disable_cull=False, coloured=False, sweep_block_worlds=0, large_scene=0, disable_islands=False,
                 sweep_form=0):
        self.L = lib()
        self.scene = scene
        cfg = BatchCfg()
        self.L.rp_batch_cfg_default(C.byref(cfg))
        cfg.max_pairs_per_world = max_pairs
        cfg.max_contacts_per_world = max_contacts
        cfg.disable_cull = int(disable_cull)
        cfg.solve_order = 1 if coloured else 0  # RP_ORDER_COLOURED / RP_ORDER_REFERENCE
        cfg.sweep_block_worlds = sweep_block_worlds  # 0: level-major sweeps
        cfg.disable_islands = int(disable_islands)
        cfg.sweep_form = int(sweep_form)  # 0 choose, 1 grid barriers between levels, 2 dataflow (rawphys_b200.h)
        cfg.large_scene = large_scene  # 0: grid broadphase / union-find islands / parallel colouring from 4096 bodies per world; 1 never; 2 always
        h = C.c_void_p()
        _check(self.L.rp_batch_create(scene.h, n_worlds, device, C.byref(cfg), C.byref(h)), "rp_batch_create")
        self.h = h
        self.W = int(self.L.rp_batch_num_worlds(h))
        self.NB = int(self.L.rp_batch_num_bodies(h))

    @classmethod
    def create_from(cls, scene, src, new_from_old):
        """A batch for `scene` whose body i continues body new_from_old[i] of `src` in every world (-1: a new body, from the
        scene's initial state): entity counts that change mid-run (rp_batch_create_from)."""
        self = cls.__new__(cls)
        self.L = lib()
        self.scene = scene
        m = np.ascontiguousarray(new_from_old, dtype=np.int32)
        if m.shape != (scene.n,):
            raise ValueError("one entry per body of the new scene")
        h = C.c_void_p()
        _check(self.L.rp_batch_create_from(scene.h, src.h, m.ctypes.data_as(_i32p), None, C.byref(h)), "rp_batch_create_from")
        self.h = h
        self.W = int(self.L.rp_batch_num_worlds(h))
        self.NB = int(self.L.rp_batch_num_bodies(h))
        return self

    def close(self):
        if getattr(self, "h", None):
            self.L.rp_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # forces (shared by all worlds), as the examples' update() adds them before every pbd_simulate call
    def clear_forces(self):
        _check(self.L.rp_batch_clear_forces(self.h), "rp_batch_clear_forces")

    def add_force(self, body, position, force):
        _check(self.L.rp_batch_add_force(self.h, body, _d(_vec(position)), _d(_vec(force))), "rp_batch_add_force")

    def add_gravity(self, g=10.0):
        _check(self.L.rp_batch_add_gravity(self.h, g), "rp_batch_add_gravity")

    def set_scene_forces(self, desc):
        """gravity + persistent forces of a scene description (what ref_step re-adds every frame)."""
        self.clear_forces()
        if desc.gravity is not None:
            self.add_gravity(desc.gravity)
        for (bi, p, f) in desc.forces:
            self.add_force(bi, p, f)

    def step(self, dt=1.0 / 60.0, substeps=20, iters=1, collisions=True):
        _check(self.L.rp_batch_step(self.h, dt, substeps, iters, int(collisions)), "rp_batch_step")

    def sync(self):
        _check(self.L.rp_batch_sync(self.h), "rp_batch_sync")

    def run(self, frames, dt=1.0 / 60.0, substeps=20, iters=1, collisions=True):
        ms = C.c_float()
        _check(self.L.rp_batch_run(self.h, frames, dt, substeps, iters, int(collisions), C.byref(ms)), "rp_batch_run")
        return float(ms.value)

    def state(self, first=0, n=None, ignore_capacity=False):
        """[n, NB, STATE_STRIDE] records. Raises RawPhysCapacityError if some world ran out of pair / contact capacity since the
        last clear_status() (its trajectory no longer follows the reference) unless `ignore_capacity`."""
        n = self.W - first if n is None else n
        out = np.zeros((n, self.NB, STATE_STRIDE))
        rc = self.L.rp_batch_download_state(self.h, first, n, out.ctypes.data)
        if not (rc == 3 and ignore_capacity):
            _check(rc, "rp_batch_download_state")
        return out

    def graph_kernels(self):
        """kernel launches inside the CUDA graph of one step()"""
        return int(self.L.rp_batch_graph_kernels(self.h))

    def clear_status(self):
        _check(self.L.rp_batch_clear_status(self.h), "rp_batch_clear_status")

    def upload(self, state, first=0):
        st = np.ascontiguousarray(state, dtype=np.float64).reshape(-1, self.NB, STATE_STRIDE)
        _check(self.L.rp_batch_upload_state(self.h, first, st.shape[0], st.ctypes.data), "rp_batch_upload_state")

    def broadcast(self, state_one_world):
        st = np.ascontiguousarray(state_one_world, dtype=np.float64).reshape(self.NB, STATE_STRIDE)
        _check(self.L.rp_batch_broadcast_state(self.h, st.ctypes.data), "rp_batch_broadcast_state")

    def step_host(self, ptr_in, ptr_out, dt=1.0 / 60.0, substeps=20, iters=1, collisions=True):
        """Host-buffer step: raw addresses of [W][NB][21] double arrays (pinned for full PCIe speed)."""
        _check(self.L.rp_batch_step_host(self.h, ptr_in, ptr_out, dt, substeps, iters, int(collisions)), "rp_batch_step_host")

    def status(self):
        out = np.zeros(self.W, dtype=np.int32)
        _check(self.L.rp_batch_get_status(self.h, out.ctypes.data_as(_i32p)), "rp_batch_get_status")
        return out

    def counters(self):
        out = np.zeros(8, dtype=np.uint64)
        _check(self.L.rp_batch_get_counters(self.h, out.ctypes.data_as(_u64p)), "rp_batch_get_counters")
        return dict(pair_tests=int(out[0]), gjk_hits=int(out[1]), contacts=int(out[2]), broad_pairs=int(out[3]), levels=int(out[4]),
                    frames=int(out[5]), gjk_runs=int(out[6]))

    def profile(self, frames, dt=1.0 / 60.0, substeps=20, iters=1, collisions=True):
        """device ms per kernel family over `frames` un-graphed frames"""
        out = (C.c_float * len(KERNEL_FAMILIES))()
        _check(self.L.rp_batch_profile(self.h, frames, dt, substeps, iters, int(collisions), out), "rp_batch_profile")
        return {k: float(out[i]) for i, k in enumerate(KERNEL_FAMILIES)}

    def step_logged(self, world=0, dt=1.0 / 60.0, substeps=20, iters=1, collisions=True, max_calls=1 << 16, max_contacts=1 << 18):
        calls = np.zeros((max_calls, 4), dtype=np.uint32)
        contacts = np.zeros((max_contacts, 9))
        nc, nk = C.c_uint32(), C.c_uint32()
        _check(self.L.rp_batch_step_logged(self.h, dt, substeps, iters, int(collisions), world, _u(calls), max_calls, _d(contacts),
                                           max_contacts, C.byref(nc), C.byref(nk)), "rp_batch_step_logged")
        if nc.value > max_calls or nk.value > max_contacts:
            raise RawPhysError("contact log truncated")
        return calls[:nc.value].copy(), contacts[:nk.value].copy()

    def broad_pairs(self, world=0, max_pairs=1 << 20):
        buf = np.zeros((max_pairs, 2), dtype=np.uint32)
        n = C.c_uint32()
        _check(self.L.rp_batch_broad_pairs(self.h, world, _u(buf), max_pairs, C.byref(n)), "rp_batch_broad_pairs")
        return buf[:n.value].astype(np.int64)


def _pair_levels(self, world=0, max_pairs=1 << 21):
    """(pairs [n, 2], levels [n]) of one world's schedule for its current poses"""
    pairs = np.zeros((max_pairs, 2), dtype=np.uint32)
    levels = np.zeros(max_pairs, dtype=np.int32)
    n = C.c_uint32()
    _check(self.L.rp_batch_pair_levels(self.h, world, _u(pairs), levels.ctypes.data_as(_i32p), max_pairs, C.byref(n)), "rp_batch_pair_levels")
    return pairs[:n.value].astype(np.int64), levels[:n.value].copy()


Batch.pair_levels = _pair_levels


def measure_fp64_peak(device=0):
    """(no-FMA TFLOP/s, FMA TFLOP/s) of the FP64 CUDA-core pipe, measured on the device"""
    out = np.zeros(2)
    _check(lib().rp_measure_fp64_peak(device, _d(out)), "rp_measure_fp64_peak")
    return float(out[0]), float(out[1])


def state15_to_21(st15):
    """refdrv / scenes 15-double records -> RP_STATE_STRIDE records (previous velocities zero, as entity_create leaves them)."""
    st15 = np.asarray(st15, dtype=np.float64)
    out = np.zeros(st15.shape[:-1] + (STATE_STRIDE,))
    out[..., :15] = st15
    return out
