"""ctypes binding of oracle/_ref/libref_oracle.so (the UNMODIFIED reference compiled by oracle/Makefile).

TEST INFRASTRUCTURE ONLY. The library is prebuilt in the build container (it needs /root/reference to compile) and
travels to the GPU box as a binary; nothing here reads /root/reference at run time.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STATE_STRIDE = 15
PARAM_STRIDE = 25

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)


def _d(a):
    return a.ctypes.data_as(_dp)


def _u(a):
    return a.ctypes.data_as(_u32p)


def lib_path(flavour="strict"):
    if flavour == "port":
        return os.path.join(ROOT, "oracle", "libport_oracle.so")
    # the reference / the restatement built WITHOUT USE_QUATERNIONS_LINEARIZED_FORMULAS (oracle/Makefile ref_exactq, port_exactq)
    if flavour == "port_exactq":
        return os.path.join(ROOT, "oracle", "libport_oracle_exactq.so")
    if flavour == "strict_exactq":
        return os.path.join(ROOT, "oracle", "_ref", "libref_oracle_exactq.so")
    # "shim": the reference's own object code for everything except the frame step, whose two entry points are
    # redirected to raw-physics_b200/shim/pbd_b200.cpp -> librawphys_b200.so (oracle/Makefile `shim`); needs a GPU to step
    name = {"strict": "libref_oracle.so", "shim": "libref_shim.so"}.get(flavour, "libref_oracle_fastmath.so")
    return os.path.join(ROOT, "oracle", "_ref", name)


class _Prefixed:
    """Resolves L.ref_xyz to <prefix>_xyz so the reference driver and the CPU restatement (oracle/port, exported as
    port_*) are driven through the very same code."""

    def __init__(self, lib, prefix):
        self._lib, self._prefix = lib, prefix

    def __getattr__(self, name):
        if name.startswith("ref_"):
            name = self._prefix + name[3:]
        return getattr(self._lib, name)


def available(flavour="strict"):
    return os.path.exists(lib_path(flavour))


class RefWorld:
    """One reference world (the reference keeps its entity table in process globals, so this is a singleton)."""

    def __init__(self, flavour="strict"):
        self.flavour = flavour
        self.lib = _Prefixed(C.CDLL(lib_path(flavour)), "port" if flavour.startswith("port") else "ref")
        L = self.lib
        L.ref_entity_create.restype = C.c_uint64
        L.ref_entity_create.argtypes = [_dp, _dp, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double]
        L.ref_collider_add_hull.argtypes = [_dp, C.c_uint32, _u32p, C.c_uint32]
        L.ref_collider_add_sphere.argtypes = [C.c_float]
        L.ref_step.argtypes = [C.c_double, C.c_uint32, C.c_uint32, C.c_int]
        L.ref_run_timed.restype = C.c_double
        L.ref_run_timed.argtypes = [C.c_uint32, C.c_double, C.c_uint32, C.c_uint32, C.c_int]
        L.ref_set_gravity.argtypes = [C.c_int, C.c_double]
        L.ref_add_persistent_force.argtypes = [C.c_uint32, _dp, _dp]
        L.ref_quaternion_new.argtypes = [_dp, C.c_double, _dp]
        L.ref_num_entities.restype = C.c_uint32
        L.ref_add_positional_constraint.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, C.c_double, _dp]
        L.ref_add_mutual_orientation_constraint.argtypes = [C.c_uint64, C.c_uint64, C.c_double]
        L.ref_add_hinge_constraint.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.c_double, C.c_double]
        L.ref_add_spherical_constraint.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                                   C.c_double, C.c_double, C.c_double, C.c_double]
        L.ref_probe_pair.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _dp, _dp, C.c_uint32]
        L.ref_broad_pairs.restype = C.c_uint32
        L.ref_broad_pairs.argtypes = [C.POINTER(C.c_uint64), C.c_uint32]
        L.ref_log_num_calls.restype = C.c_uint32
        L.ref_log_num_contacts.restype = C.c_uint32
        L.ref_total_narrowphase_calls.restype = C.c_uint64
        L.ref_reset()

    # ---- scene construction from a scenes.Scene
    def load(self, scene):
        L = self.lib
        L.ref_reset()
        for b in scene.bodies:
            L.ref_collider_begin()
            for col in b.colliders:
                if col.kind == "sphere":
                    L.ref_collider_add_sphere(C.c_float(col.radius))
                else:
                    v = np.ascontiguousarray(col.vertices, dtype=np.float64)
                    idx = np.ascontiguousarray(col.indices, dtype=np.uint32)
                    L.ref_collider_add_hull(_d(v), v.shape[0], _u(idx), idx.shape[0])
            pos = np.asarray(b.position, dtype=np.float64)
            quat = np.asarray(b.rotation, dtype=np.float64)
            L.ref_entity_create(_d(pos), _d(quat), b.mass, int(b.fixed), b.mu_s, b.mu_d, b.restitution)
        L.ref_set_gravity(int(scene.gravity is not None), scene.gravity if scene.gravity is not None else 0.0)
        for (bi, p, f) in scene.forces:
            pp = np.asarray(p, dtype=np.float64)
            ff = np.asarray(f, dtype=np.float64)
            L.ref_add_persistent_force(bi, _d(pp), _d(ff))
        for c in scene.constraints:
            r1 = np.asarray(c.get("r1", (0, 0, 0)), dtype=np.float64)
            r2 = np.asarray(c.get("r2", (0, 0, 0)), dtype=np.float64)
            if c["type"] == "positional":
                dist = np.asarray(c["distance"], dtype=np.float64)
                L.ref_add_positional_constraint(c["e1"], c["e2"], _d(r1), _d(r2), c["compliance"], _d(dist))
            elif c["type"] == "mutual_orientation":
                L.ref_add_mutual_orientation_constraint(c["e1"], c["e2"], c["compliance"])
            elif c["type"] == "hinge":
                L.ref_add_hinge_constraint(c["e1"], c["e2"], _d(r1), _d(r2), c["compliance"], c["e1_aligned"], c["e2_aligned"],
                                           int(c["limited"]), c.get("e1_limit", 0), c.get("e2_limit", 0),
                                           c.get("lower", 0.0), c.get("upper", 0.0))
            elif c["type"] == "spherical":
                L.ref_add_spherical_constraint(c["e1"], c["e2"], _d(r1), _d(r2), c["e1_swing"], c["e2_swing"], c["e1_twist"],
                                               c["e2_twist"], c["swing_lower"], c["swing_upper"], c["twist_lower"],
                                               c["twist_upper"])
            else:
                raise ValueError(c["type"])
        if scene.initial_state is not None:
            self.set_state(scene.initial_state)
        self.n = int(L.ref_num_entities())
        return self

    def add_body(self, b):
        """entity_create[_fixed] for one more body of a scenes.BodyDesc kind, at any time (bodies thrown in mid-run)"""
        L = self.lib
        L.ref_collider_begin()
        for col in b.colliders:
            if col.kind == "sphere":
                L.ref_collider_add_sphere(C.c_float(col.radius))
            else:
                v = np.ascontiguousarray(col.vertices, dtype=np.float64)
                idx = np.ascontiguousarray(col.indices, dtype=np.uint32)
                L.ref_collider_add_hull(_d(v), v.shape[0], _u(idx), idx.shape[0])
        pos = np.asarray(b.position, dtype=np.float64)
        quat = np.asarray(b.rotation, dtype=np.float64)
        L.ref_entity_create(_d(pos), _d(quat), b.mass, int(b.fixed), b.mu_s, b.mu_d, b.restitution)
        self.n = int(L.ref_num_entities())

    def quaternion_new(self, axis, angle_degrees):
        a = np.asarray(axis, dtype=np.float64)
        out = np.zeros(4)
        self.lib.ref_quaternion_new(_d(a), angle_degrees, _d(out))
        return out

    def step(self, dt=1.0 / 60.0, substeps=20, iters=1, collisions=True):
        self.lib.ref_step(dt, substeps, iters, int(collisions))

    def run_timed(self, frames, dt=1.0 / 60.0, substeps=20, iters=1, collisions=True):
        return float(self.lib.ref_run_timed(frames, dt, substeps, iters, int(collisions)))

    def state(self):
        n = int(self.lib.ref_num_entities())
        out = np.zeros((n, STATE_STRIDE))
        self.lib.ref_get_state(_d(out))
        return out

    def prev_velocities(self):
        """previous_linear_velocity, previous_angular_velocity per entity (what RP_STATE_STRIDE records hold at [15:21])"""
        n = int(self.lib.ref_num_entities())
        out = np.zeros((n, 6))
        self.lib.ref_get_prev_velocities(_d(out))
        return out

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=np.float64)
        self.lib.ref_set_state(_d(st))

    def params(self):
        n = int(self.lib.ref_num_entities())
        out = np.zeros((n, PARAM_STRIDE))
        self.lib.ref_get_params(_d(out))
        return out

    def hull(self, entity, collider=0):
        sizes = np.zeros(6, dtype=np.int32)
        self.lib.ref_hull_sizes(entity, collider, sizes.ctypes.data_as(C.POINTER(C.c_int32)))
        if sizes[0] < 0:
            return None
        V, F, fe, v2f, v2n, f2n = [int(x) for x in sizes]
        h = dict(verts=np.zeros((V, 3)), normals=np.zeros((F, 3)),
                 face_ptr=np.zeros(F + 1, np.uint32), face_idx=np.zeros(fe, np.uint32),
                 v2f_ptr=np.zeros(V + 1, np.uint32), v2f_idx=np.zeros(v2f, np.uint32),
                 v2n_ptr=np.zeros(V + 1, np.uint32), v2n_idx=np.zeros(v2n, np.uint32),
                 f2n_ptr=np.zeros(F + 1, np.uint32), f2n_idx=np.zeros(f2n, np.uint32))
        self.lib.ref_hull_dump(entity, collider, _d(h["verts"]), _d(h["normals"]), _u(h["face_ptr"]), _u(h["face_idx"]),
                               _u(h["v2f_ptr"]), _u(h["v2f_idx"]), _u(h["v2n_ptr"]), _u(h["v2n_idx"]),
                               _u(h["f2n_ptr"]), _u(h["f2n_idx"]))
        return h

    def probe_pair(self, ia, ib, ca=0, cb=0, max_contacts=256):
        out = np.zeros(19)
        contacts = np.zeros((max_contacts, 9))
        self.lib.ref_probe_pair(ia, ca, ib, cb, _d(out), _d(contacts), max_contacts)
        nc = int(out[18])
        return dict(hit=bool(out[0]), simplex=out[1:13].reshape(4, 3).copy(), epa_ok=bool(out[13]),
                    normal=out[14:17].copy(), penetration=float(out[17]), contacts=contacts[:nc].copy())

    def broad_pairs(self, max_pairs=1 << 20):
        buf = np.zeros((max_pairs, 2), dtype=np.uint64)
        n = int(self.lib.ref_broad_pairs(buf.ctypes.data_as(C.POINTER(C.c_uint64)), max_pairs))
        return buf[:n].astype(np.int64)

    # ---- contact log
    def log_enable(self, on=True):
        self.lib.ref_log_enable(int(on))

    def log_clear(self):
        self.lib.ref_log_clear()

    def log_get(self):
        nc = int(self.lib.ref_log_num_calls())
        nk = int(self.lib.ref_log_num_contacts())
        calls = np.zeros((nc, 4), dtype=np.uint32)
        contacts = np.zeros((nk, 9))
        self.lib.ref_log_get(_u(calls), _d(contacts))
        return calls, contacts


def split_substeps(calls):
    """Splits the --wrap call log into substeps: pairs are visited in ascending (e1, e2) order inside a substep
    (pbd.cpp:584), so a non-increasing pair key starts a new substep."""
    out, cur, prev = [], [], None
    for row in calls:
        key = (int(row[0]), int(row[1]))
        if prev is not None and key <= prev:
            out.append(cur)
            cur = []
        cur.append(row)
        prev = key
    if cur:
        out.append(cur)
    return out
