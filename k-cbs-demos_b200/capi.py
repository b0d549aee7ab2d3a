"""ctypes binding of include/kcbs_b200.h.  No compute happens in Python and there is no fallback:
if libkcbs_b200.so is missing or a context cannot be created, the call raises."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[C.CDLL] = None

KCBS_MAX_OBSTACLES = 64
KCBS_MAX_ROBOTS = 64
KCBS_MAX_STEPS = 32
NO_CONFLICT_KEY = 0xFFFFFFFFFFFFFFFF


class KcbsError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"kcbs status {status}: {message}")
        self.status = status


class _CWorld(C.Structure):
    _fields_ = [
        ("lo", C.c_double * 4),
        ("hi", C.c_double * 4),
        ("step_size", C.c_double),
        ("integration_step", C.c_double),
        ("car_length", C.c_double),
        ("n_obstacles", C.c_int32),
        ("n_robots", C.c_int32),
        ("obstacles", C.POINTER(C.c_double)),
        ("robot_dims", C.POINTER(C.c_double)),
    ]


class Conflict(C.Structure):
    """kcbs_conflict: first conflict of a plan (r1 = -1 when conflict free)."""

    _fields_ = [("r1", C.c_int32), ("r2", C.c_int32), ("k_start", C.c_int32), ("k_end", C.c_int32)]

    def as_tuple(self):
        return (self.r1, self.r2, self.k_start, self.k_end)


class RrtParams(C.Structure):
    """kcbs_rrt_params"""

    _fields_ = [("n_targets", C.c_int32), ("n_controls", C.c_int32), ("max_nodes", C.c_int32),
                ("max_iterations", C.c_int32), ("min_steps", C.c_int32), ("max_steps", C.c_int32),
                ("goal_bias", C.c_float), ("goal_radius", C.c_float), ("time_limit_s", C.c_double),
                ("seed", C.c_uint64)]


class RrtStats(C.Structure):
    _fields_ = [("solved", C.c_int32), ("iterations", C.c_int32), ("nodes", C.c_int32), ("path_steps", C.c_int32),
                ("seconds", C.c_double)]


class KcbsParams(C.Structure):
    _fields_ = [("low_level", RrtParams), ("time_limit_s", C.c_double), ("max_nodes", C.c_int32), ("merge_bound", C.c_int32)]


class KcbsStats(C.Structure):
    _fields_ = [("solved", C.c_int32), ("nodes_expanded", C.c_int32), ("approximate_solutions", C.c_int32),
                ("low_level_calls", C.c_int32), ("root_seconds", C.c_double), ("seconds", C.c_double),
                ("conflict_checks", C.c_int64), ("merges", C.c_int32), ("reserved_", C.c_int32)]


CONFLICT_DTYPE = np.dtype([("r1", "<i4"), ("r2", "<i4"), ("k_start", "<i4"), ("k_end", "<i4")])

EXPORTS = [
    "kcbs_abi_version", "kcbs_status_string", "kcbs_world_defaults", "kcbs_create", "kcbs_destroy",
    "kcbs_last_error", "kcbs_sync", "kcbs_host_alloc", "kcbs_host_free", "kcbs_propagate_batch",
    "kcbs_propagate_batch_dev", "kcbs_validity_batch", "kcbs_validity_batch_dev", "kcbs_first_conflict",
    "kcbs_first_conflict_dev", "kcbs_conflict_keys_dev", "kcbs_conflict_finish_dev",
    "kcbs_measure_fp32_peak", "kcbs_measure_fp32x2_peak", "kcbs_launch_count", "kcbs_rrt_defaults",
    "kcbs_kcbs_defaults", "kcbs_rrt_plan", "kcbs_kcbs_solve", "kcbs_propagate_pair_batch", "kcbs_propagate_pair_batch_dev",
    "kcbs_validity_pair_batch", "kcbs_validity_pair_batch_dev", "kcbs_first_conflict_groups", "kcbs_first_conflict_groups_dev",
    "kcbs_scene_grid", "kcbs_rrt_plan_many", "kcbs_peer_export", "kcbs_peer_connect", "kcbs_peer_disconnect",
    "kcbs_peer_status", "kcbs_time_slab", "kcbs_first_conflict_sharded_dev", "kcbs_propagate_packed",
    "kcbs_propagate_packed_dev", "kcbs_set_expansion_mode", "kcbs_rrt_plan_pair",
]
KCBS_MAX_PEERS = 16
PEER_HANDLE_BYTES = 128


def lib_path() -> str:
    # KCBS_B200_LIB selects an alternative build of the same library (A/B experiments); never a different backend
    return os.environ.get("KCBS_B200_LIB") or os.path.join(_HERE, "libkcbs_b200.so")


def load_library() -> C.CDLL:
    """Loads libkcbs_b200.so (built in-tree by __graft_entry__.build()); raises if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no Python/CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.kcbs_abi_version.restype = C.c_int
    lib.kcbs_status_string.restype = C.c_char_p
    lib.kcbs_status_string.argtypes = [C.c_int]
    lib.kcbs_world_defaults.restype = None
    lib.kcbs_world_defaults.argtypes = [C.POINTER(_CWorld), C.c_double, C.c_double]
    lib.kcbs_create.argtypes = [C.POINTER(_CWorld), i32, C.POINTER(vp)]
    lib.kcbs_destroy.restype = None
    lib.kcbs_destroy.argtypes = [vp]
    lib.kcbs_last_error.restype = C.c_char_p
    lib.kcbs_last_error.argtypes = [vp]
    lib.kcbs_sync.argtypes = [vp]
    lib.kcbs_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.kcbs_host_free.argtypes = [vp, vp]
    prop = [vp, i32, i64, i64, vp, vp, vp, vp, i32, vp, vp, vp]
    lib.kcbs_propagate_batch.argtypes = prop
    lib.kcbs_propagate_batch_dev.argtypes = prop + [vp]
    packed = [vp, i32, i64, i64, vp, vp, vp, vp, i32, vp, i64, vp, vp, vp]
    lib.kcbs_propagate_packed.argtypes = packed + [C.POINTER(C.c_int64)]
    lib.kcbs_propagate_packed_dev.argtypes = packed + [vp]
    lib.kcbs_set_expansion_mode.argtypes = [vp, i32]
    lib.kcbs_validity_batch.argtypes = [vp, i32, i64, vp, vp]
    lib.kcbs_validity_batch_dev.argtypes = [vp, i32, i64, vp, vp, vp]
    lib.kcbs_first_conflict.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    lib.kcbs_first_conflict_dev.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    lib.kcbs_conflict_keys_dev.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.kcbs_conflict_finish_dev.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.kcbs_measure_fp32_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.kcbs_measure_fp32x2_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.kcbs_rrt_defaults.restype = None
    lib.kcbs_rrt_defaults.argtypes = [C.POINTER(RrtParams)]
    lib.kcbs_kcbs_defaults.restype = None
    lib.kcbs_kcbs_defaults.argtypes = [C.POINTER(KcbsParams)]
    lib.kcbs_rrt_plan.argtypes = [vp, i32, vp, vp, C.POINTER(RrtParams), i32, vp, vp, vp, vp, vp, i32, vp,
                                  C.POINTER(C.c_int32), C.POINTER(RrtStats)]
    lib.kcbs_rrt_plan_pair.argtypes = [vp, i32, i32, vp, vp, C.POINTER(RrtParams), i32, vp, vp, vp, vp, vp, i32, vp,
                                       C.POINTER(C.c_int32), C.POINTER(RrtStats)]
    lib.kcbs_rrt_plan_many.argtypes = [vp, i32, vp, vp, vp, C.POINTER(RrtParams), vp, i32, vp, vp, vp]
    lib.kcbs_kcbs_solve.argtypes = [vp, i32, vp, vp, C.POINTER(KcbsParams), i32, vp, vp, C.POINTER(KcbsStats)]
    pair = [vp, i32, i32, i64, i64, vp, vp, vp, vp, i32, vp, C.c_float, vp, vp, vp]
    lib.kcbs_propagate_pair_batch.argtypes = pair
    lib.kcbs_propagate_pair_batch_dev.argtypes = pair + [vp]
    lib.kcbs_validity_pair_batch.argtypes = [vp, i32, i32, i64, vp, vp]
    lib.kcbs_validity_pair_batch_dev.argtypes = [vp, i32, i32, i64, vp, vp, vp]
    lib.kcbs_first_conflict_groups.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    lib.kcbs_first_conflict_groups_dev.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.kcbs_scene_grid.argtypes = [C.POINTER(_CWorld), C.POINTER(C.c_int32), vp]
    lib.kcbs_peer_export.argtypes = [vp, i32, vp]
    lib.kcbs_peer_connect.argtypes = [vp, i32, i32, vp]
    lib.kcbs_peer_disconnect.argtypes = [vp]
    lib.kcbs_peer_status.argtypes = [vp]
    lib.kcbs_time_slab.argtypes = [i32, i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.kcbs_first_conflict_sharded_dev.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    lib.kcbs_launch_count.restype = i64
    lib.kcbs_launch_count.argtypes = [vp]
    _LIB = lib
    return lib


@dataclass
class World:
    """Scene description (kcbs_world).  obstacles: rows (x, y, length, width) as given to
    RectangularObstacle (Obstacle.h:82); robot_dims: rows (length, width) as given to
    RectangularRobot (Robot.h:82), in addIndividual order."""

    x_max: float
    y_max: float
    robot_dims: Sequence[Sequence[float]]
    obstacles: Sequence[Sequence[float]] = field(default_factory=list)
    lo: Optional[Sequence[float]] = None
    hi: Optional[Sequence[float]] = None
    step_size: float = 0.1
    integration_step: float = 1e-2
    car_length: float = 0.5

    def bounds(self):
        lo = list(self.lo) if self.lo is not None else [0.0, 0.0, -1.0, -np.pi / 3]
        hi = list(self.hi) if self.hi is not None else [float(self.x_max), float(self.y_max), 1.0, np.pi / 3]
        return lo, hi


def _stream(s) -> Optional[int]:
    """cudaStream_t for the *_dev calls.  None -> the context's own stream.  torch reports its default
    stream as handle 0, which the C ABI would read as "context stream"; map it to cudaStreamLegacy."""
    if s is None:
        return None
    if hasattr(s, "cuda_stream"):
        s = s.cuda_stream
    return 1 if int(s) == 0 else int(s)


def _ptr(a) -> Optional[int]:
    """Host pointer of a C-contiguous numpy array, device pointer of a torch tensor, or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous(), "tensor must be contiguous"
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError(type(a))


def _c_world(world: World):
    """kcbs_world for `world`, plus the arrays it points into (keep them alive while it is used)."""
    lo, hi = world.bounds()
    obs = np.ascontiguousarray(np.asarray(world.obstacles, dtype=np.float64).reshape(-1, 4))
    dims = np.ascontiguousarray(np.asarray(world.robot_dims, dtype=np.float64).reshape(-1, 2))
    cw = _CWorld()
    cw.lo[:] = lo
    cw.hi[:] = hi
    cw.step_size, cw.integration_step, cw.car_length = world.step_size, world.integration_step, world.car_length
    cw.n_obstacles, cw.n_robots = len(obs), len(dims)
    cw.obstacles = obs.ctypes.data_as(C.POINTER(C.c_double)) if len(obs) else None
    cw.robot_dims = dims.ctypes.data_as(C.POINTER(C.c_double))
    return cw, obs, dims


def scene_grid(world: World) -> np.ndarray:
    """kcbs_scene_grid: the [dim, dim] uint16 obstacle cell codes derived from `world` (host only, no device needed)."""
    lib = load_library()
    cw, _obs, _dims = _c_world(world)
    dim = C.c_int32()
    st = lib.kcbs_scene_grid(C.byref(cw), C.byref(dim), None)
    if st != 0:
        raise KcbsError(st, lib.kcbs_last_error(None).decode())
    codes = np.zeros((dim.value, dim.value), np.uint16)
    st = lib.kcbs_scene_grid(C.byref(cw), C.byref(dim), codes.ctypes.data)
    if st != 0:
        raise KcbsError(st, lib.kcbs_last_error(None).decode())
    return codes


class Context:
    """kcbs_ctx.  Host-array methods mirror the reference-facing host entry points; *_dev methods
    take torch CUDA tensors (or raw device pointers) and are asynchronous on `stream`."""

    def __init__(self, world: World, device: int = 0):
        self._lib = load_library()
        self.world = world
        cw, self._obs, self._dims = _c_world(world)
        h = C.c_void_p()
        st = self._lib.kcbs_create(C.byref(cw), device, C.byref(h))
        if st != 0:
            raise KcbsError(st, self._lib.kcbs_last_error(None).decode())
        self._h = h
        self.device = device

    # -- plumbing -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            for p in getattr(self, "_pinned", []):
                self._lib.kcbs_host_free(self._h, p)
            self._pinned = []
            self._lib.kcbs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, st: int):
        if st != 0:
            raise KcbsError(st, self._lib.kcbs_last_error(self._h).decode())

    def sync(self):
        self._check(self._lib.kcbs_sync(self._h))

    def launch_count(self) -> int:
        return int(self._lib.kcbs_launch_count(self._h))

    def measure_fp32_peak(self) -> float:
        v = C.c_double()
        self._check(self._lib.kcbs_measure_fp32_peak(self._h, C.byref(v)))
        return v.value

    def measure_fp32x2_peak(self) -> float:
        v = C.c_double()
        self._check(self._lib.kcbs_measure_fp32x2_peak(self._h, C.byref(v)))
        return v.value

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array over pinned host memory from kcbs_host_alloc (freed with the context)."""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._check(self._lib.kcbs_host_alloc(self._h, nbytes, C.byref(p)))
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        return arr

    # -- (1) propagation ------------------------------------------------------------------
    def propagate_batch(self, robot, parents, controls, steps=None, parent_idx=None, max_steps=10,
                        want_seg=True, want_final=True, out=None):
        """Host entry point kcbs_propagate_batch.  parents [5, n_parents], controls [2, n] float32.
        Returns (seg [5, max_steps, n] or None, final [5, n] or None, valid_steps [n])."""
        parents = np.ascontiguousarray(parents, dtype=np.float32)
        controls = np.ascontiguousarray(controls, dtype=np.float32)
        n, n_parents = controls.shape[1], parents.shape[1]
        steps = None if steps is None else np.ascontiguousarray(steps, dtype=np.int32)
        parent_idx = None if parent_idx is None else np.ascontiguousarray(parent_idx, dtype=np.int32)
        if out is not None:
            seg, fin, valid = out
        else:
            seg = np.zeros((5, max_steps, n), np.float32) if want_seg else None
            fin = np.zeros((5, n), np.float32) if want_final else None
            valid = np.zeros(n, np.int32)
        self._check(self._lib.kcbs_propagate_batch(self._h, robot, n, n_parents, _ptr(parents), _ptr(parent_idx),
                                                   _ptr(controls), _ptr(steps), max_steps, _ptr(seg), _ptr(fin),
                                                   _ptr(valid)))
        return seg, fin, valid

    def propagate_batch_dev(self, robot, n, n_parents, parents, controls, steps, parent_idx, max_steps, seg, fin,
                            valid, stream=None):
        self._check(self._lib.kcbs_propagate_batch_dev(self._h, robot, n, n_parents, _ptr(parents), _ptr(parent_idx),
                                                       _ptr(controls), _ptr(steps), max_steps, _ptr(seg), _ptr(fin),
                                                       _ptr(valid), _stream(stream)))

    def propagate_packed(self, robot, parents, controls, steps=None, parent_idx=None, max_steps=10, want_final=True,
                         out=None, packed_cap=None):
        """Host entry point kcbs_propagate_packed.  Returns (packed [5, cap] (rows [:, :total] are in use), offsets [n + 1],
        final [5, n] or None, valid_steps [n], total).  Sample i's states are rows offsets[i] .. offsets[i] + valid_steps[i] - 1;
        blocks of samples are placed first come, first served, so offsets is not monotone across blocks (offsets[n] = total).
        out = (packed, offsets, final, valid) reuses caller buffers."""
        parents = np.ascontiguousarray(parents, dtype=np.float32)
        controls = np.ascontiguousarray(controls, dtype=np.float32)
        n, n_parents = controls.shape[1], parents.shape[1]
        steps = None if steps is None else np.ascontiguousarray(steps, dtype=np.int32)
        parent_idx = None if parent_idx is None else np.ascontiguousarray(parent_idx, dtype=np.int32)
        if out is not None:
            packed, offsets, fin, valid = out
        else:
            cap = n * max_steps if packed_cap is None else packed_cap
            packed = np.zeros((5, cap), np.float32)
            offsets = np.zeros(n + 1, np.int32)
            fin = np.zeros((5, n), np.float32) if want_final else None
            valid = np.zeros(n, np.int32)
        total = C.c_int64()
        self._check(self._lib.kcbs_propagate_packed(self._h, robot, n, n_parents, _ptr(parents), _ptr(parent_idx),
                                                    _ptr(controls), _ptr(steps), max_steps, _ptr(packed), packed.shape[1],
                                                    _ptr(offsets), _ptr(fin), _ptr(valid), C.byref(total)))
        return packed, offsets, fin, valid, total.value

    def propagate_packed_dev(self, robot, n, n_parents, parents, controls, steps, parent_idx, max_steps, packed, packed_cap,
                             offsets, fin, valid, stream=None):
        self._check(self._lib.kcbs_propagate_packed_dev(self._h, robot, n, n_parents, _ptr(parents), _ptr(parent_idx),
                                                        _ptr(controls), _ptr(steps), max_steps, _ptr(packed), packed_cap,
                                                        _ptr(offsets), _ptr(fin), _ptr(valid), _stream(stream)))

    def set_expansion_mode(self, mode: int):
        """kcbs_set_expansion_mode: 0 = launch shape from n, 1 = latency shape, 2 = throughput shape."""
        self._check(self._lib.kcbs_set_expansion_mode(self._h, mode))

    # -- (2) validity ---------------------------------------------------------------------
    def validity_batch(self, robot, states, out=None):
        """Host entry point kcbs_validity_batch.  states [5, n] float32 -> packed uint32 bitmask."""
        states = np.ascontiguousarray(states, dtype=np.float32)
        n = states.shape[1]
        mask = out if out is not None else np.zeros((n + 31) // 32, np.uint32)
        self._check(self._lib.kcbs_validity_batch(self._h, robot, n, _ptr(states), _ptr(mask)))
        return mask

    def validity_batch_dev(self, robot, n, states, bitmask, stream=None):
        self._check(self._lib.kcbs_validity_batch_dev(self._h, robot, n, _ptr(states), _ptr(bitmask), _stream(stream)))

    @staticmethod
    def unpack_mask(mask: np.ndarray, n: int) -> np.ndarray:
        bits = np.unpackbits(mask.view(np.uint8), bitorder="little")
        return bits[:n].astype(bool)

    # -- (3) plan validation --------------------------------------------------------------
    def first_conflict(self, traj, lengths=None, out=None):
        """Host entry point kcbs_first_conflict.  traj [P, R, 3, T] float32 (or [R, 3, T])."""
        traj = np.ascontiguousarray(traj, dtype=np.float32)
        if traj.ndim == 3:
            traj = traj[None]
        P, R, _, T = traj.shape
        lengths = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32).reshape(P, R)
        res = out if out is not None else np.zeros(P, CONFLICT_DTYPE)
        self._check(self._lib.kcbs_first_conflict(self._h, P, R, T, _ptr(traj), _ptr(lengths), _ptr(res)))
        return res

    def first_conflict_dev(self, P, R, T, traj, lengths, out, stream=None):
        self._check(self._lib.kcbs_first_conflict_dev(self._h, P, R, T, _ptr(traj), _ptr(lengths), _ptr(out), _stream(stream)))

    def conflict_keys_dev(self, P, R, T, k_lo, k_hi, traj, lengths, keys, stream=None):
        self._check(self._lib.kcbs_conflict_keys_dev(self._h, P, R, T, k_lo, k_hi, _ptr(traj), _ptr(lengths),
                                                     _ptr(keys), _stream(stream)))

    def conflict_finish_dev(self, P, R, T, traj, lengths, keys, out, stream=None):
        self._check(self._lib.kcbs_conflict_finish_dev(self._h, P, R, T, _ptr(traj), _ptr(lengths), _ptr(keys),
                                                       _ptr(out), _stream(stream)))

    # -- (3c) plan validation sharded over the GPUs of a node ---------------------------------------
    def peer_export(self, max_plans: int) -> bytes:
        """kcbs_peer_export: allocates this context's key mailbox; returns the opaque 128-byte handle."""
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        self._check(self._lib.kcbs_peer_export(self._h, max_plans, C.cast(buf, C.c_void_p)))
        return buf.raw

    def peer_connect(self, rank: int, world: int, handles: Sequence[bytes]):
        """kcbs_peer_connect with the handles of all ranks in rank order."""
        assert len(handles) == world and all(len(h) == PEER_HANDLE_BYTES for h in handles)
        buf = C.create_string_buffer(b"".join(handles), PEER_HANDLE_BYTES * world)
        self._check(self._lib.kcbs_peer_connect(self._h, rank, world, C.cast(buf, C.c_void_p)))

    def peer_disconnect(self):
        self._check(self._lib.kcbs_peer_disconnect(self._h))

    def peer_status(self):
        """Raises KcbsError (status 6) when a sharded scan since the last call gave up waiting for a peer."""
        self._check(self._lib.kcbs_peer_status(self._h))

    def time_slab(self, T: int, rank: int, world: int):
        lo, hi = C.c_int32(), C.c_int32()
        st = self._lib.kcbs_time_slab(T, rank, world, C.byref(lo), C.byref(hi))
        if st != 0:
            raise KcbsError(st, "kcbs_time_slab: invalid argument")
        return lo.value, hi.value

    def first_conflict_sharded_dev(self, P, R, T, traj, lengths, out, stream=None):
        self._check(self._lib.kcbs_first_conflict_sharded_dev(self._h, P, R, T, _ptr(traj), _ptr(lengths), _ptr(out),
                                                              _stream(stream)))

    # -- (3b) merged pair ------------------------------------------------------------------------
    def propagate_pair_batch(self, robot1, robot2, parents, controls, goals, steps=None, parent_idx=None, max_steps=10,
                             hold_radius=1.5):
        """kcbs_propagate_pair_batch.  parents [10, n_parents], controls [4, n], goals (gx1, gy1, gx2, gy2).
        Returns (seg [10, max_steps, n], final [10, n], valid_steps [n])."""
        parents = np.ascontiguousarray(parents, dtype=np.float32)
        controls = np.ascontiguousarray(controls, dtype=np.float32)
        goals = np.ascontiguousarray(goals, dtype=np.float32)
        n, n_parents = controls.shape[1], parents.shape[1]
        steps = None if steps is None else np.ascontiguousarray(steps, dtype=np.int32)
        parent_idx = None if parent_idx is None else np.ascontiguousarray(parent_idx, dtype=np.int32)
        seg = np.zeros((10, max_steps, n), np.float32)
        fin = np.zeros((10, n), np.float32)
        valid = np.zeros(n, np.int32)
        self._check(self._lib.kcbs_propagate_pair_batch(self._h, robot1, robot2, n, n_parents, _ptr(parents), _ptr(parent_idx),
                                                        _ptr(controls), _ptr(steps), max_steps, _ptr(goals), hold_radius,
                                                        _ptr(seg), _ptr(fin), _ptr(valid)))
        return seg, fin, valid

    def validity_pair_batch(self, robot1, robot2, states):
        """kcbs_validity_pair_batch.  states [10, n] -> packed uint32 bitmask."""
        states = np.ascontiguousarray(states, dtype=np.float32)
        n = states.shape[1]
        mask = np.zeros((n + 31) // 32, np.uint32)
        self._check(self._lib.kcbs_validity_pair_batch(self._h, robot1, robot2, n, _ptr(states), _ptr(mask)))
        return mask

    def first_conflict_groups(self, traj, group, lengths=None):
        """kcbs_first_conflict_groups.  traj [P, R, 3, T] rows are bodies; group [R] = individual of each body."""
        traj = np.ascontiguousarray(traj, dtype=np.float32)
        if traj.ndim == 3:
            traj = traj[None]
        P, R, _, T = traj.shape
        group = np.ascontiguousarray(group, dtype=np.int32)
        lengths = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32).reshape(P, R)
        res = np.zeros(P, CONFLICT_DTYPE)
        self._check(self._lib.kcbs_first_conflict_groups(self._h, P, R, T, _ptr(traj), _ptr(lengths), _ptr(group), _ptr(res)))
        return res

    # -- (4) next rows: batched low-level planner and K-CBS -----------------------------------
    def rrt_params(self, **kw) -> RrtParams:
        p = RrtParams()
        self._lib.kcbs_rrt_defaults(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def kcbs_params(self, low_level=None, **kw) -> KcbsParams:
        p = KcbsParams()
        self._lib.kcbs_kcbs_defaults(C.byref(p))
        if low_level is not None:
            p.low_level = low_level
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def rrt_plan(self, robot, start, goal_xy, params=None, constraints=(), max_T=4096, persistent=None):
        """kcbs_rrt_plan.  constraints: iterable of (other_robot, k0, states [3, len]); persistent[i] marks constraint i
        as holding its last state for ever (passed as a negative length).  Returns (traj [5, T] float32, RrtStats)."""
        start = np.ascontiguousarray(start, dtype=np.float32)
        goal = np.ascontiguousarray(goal_xy, dtype=np.float32)
        params = params if params is not None else self.rrt_params()
        cons = list(constraints)
        total = sum(c[2].shape[1] for c in cons)
        rob = np.array([c[0] for c in cons], np.int32)
        k0 = np.array([c[1] for c in cons], np.int32)
        ln = np.array([c[2].shape[1] for c in cons], np.int32)
        off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.int32) if cons else np.zeros(0, np.int32)
        states = np.zeros((3, max(total, 1)), np.float32)
        for c, o in zip(cons, off):
            states[:, o:o + c[2].shape[1]] = c[2]
        if persistent is not None:
            ln = np.where(np.asarray(persistent, bool), -ln, ln).astype(np.int32)
        out = np.zeros((5, max_T), np.float32)
        T = C.c_int32()
        st = RrtStats()
        self._check(self._lib.kcbs_rrt_plan(self._h, robot, _ptr(start), _ptr(goal), C.byref(params), len(cons), _ptr(rob),
                                            _ptr(k0), _ptr(ln), _ptr(off), _ptr(states), max_T, _ptr(out), C.byref(T),
                                            C.byref(st)))
        return out[:, :T.value].copy(), st

    def rrt_plan_pair(self, robot1, robot2, start10, goals_xy, params=None, constraints=(), max_T=4096, persistent=None):
        """kcbs_rrt_plan_pair: the merged pair (robot1, robot2) planned as one 10-d system.  start10 = 10 reals,
        goals_xy = (gx1, gy1, gx2, gy2).  Returns (traj [10, T] float32, RrtStats)."""
        start = np.ascontiguousarray(start10, dtype=np.float32)
        goal = np.ascontiguousarray(goals_xy, dtype=np.float32)
        assert start.size == 10 and goal.size == 4
        params = params if params is not None else self.rrt_params()
        cons = list(constraints)
        total = sum(c[2].shape[1] for c in cons)
        rob = np.array([c[0] for c in cons], np.int32)
        k0 = np.array([c[1] for c in cons], np.int32)
        ln = np.array([c[2].shape[1] for c in cons], np.int32)
        off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.int32) if cons else np.zeros(0, np.int32)
        states = np.zeros((3, max(total, 1)), np.float32)
        for c, o in zip(cons, off):
            states[:, o:o + c[2].shape[1]] = c[2]
        if persistent is not None:
            ln = np.where(np.asarray(persistent, bool), -ln, ln).astype(np.int32)
        out = np.zeros((10, max_T), np.float32)
        T = C.c_int32()
        st = RrtStats()
        self._check(self._lib.kcbs_rrt_plan_pair(self._h, robot1, robot2, _ptr(start), _ptr(goal), C.byref(params), len(cons),
                                                 _ptr(rob), _ptr(k0), _ptr(ln), _ptr(off), _ptr(states), max_T, _ptr(out),
                                                 C.byref(T), C.byref(st)))
        return out[:, :T.value].copy(), st

    def rrt_plan_many(self, robots, starts, goals, params=None, seeds=None, max_T=4096):
        """kcbs_rrt_plan_many.  robots [n], starts [5, n], goals [n, 2], seeds [n] or None.  Returns (traj [n, 5, max_T],
        lengths [n] int32, list of RrtStats): the n independent plans run concurrently on the context's planner lanes."""
        robots = np.ascontiguousarray(robots, dtype=np.int32)
        n = len(robots)
        starts = np.ascontiguousarray(starts, dtype=np.float32).reshape(5, n)
        goals = np.ascontiguousarray(goals, dtype=np.float32).reshape(n, 2)
        seeds = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        params = params if params is not None else self.rrt_params()
        traj = np.zeros((n, 5, max_T), np.float32)
        T = np.zeros(n, np.int32)
        stats = (RrtStats * max(n, 1))()
        self._check(self._lib.kcbs_rrt_plan_many(self._h, n, _ptr(robots), _ptr(starts), _ptr(goals), C.byref(params), _ptr(seeds),
                                                 max_T, _ptr(traj), _ptr(T), C.cast(stats, C.c_void_p)))
        return traj, T, list(stats)[:n]

    def kcbs_solve(self, starts, goals, params=None, max_T=4096):
        """kcbs_kcbs_solve.  starts [5, R], goals [R, 2].  Returns (plan [R, 5, max_T], lengths [R], KcbsStats)."""
        starts = np.ascontiguousarray(starts, dtype=np.float32)
        goals = np.ascontiguousarray(goals, dtype=np.float32)
        R = starts.shape[1]
        params = params if params is not None else self.kcbs_params()
        plan = np.zeros((R, 5, max_T), np.float32)
        lengths = np.zeros(R, np.int32)
        st = KcbsStats()
        self._check(self._lib.kcbs_kcbs_solve(self._h, R, _ptr(starts), _ptr(goals), C.byref(params), max_T, _ptr(plan),
                                              _ptr(lengths), C.byref(st)))
        return plan, lengths, st
