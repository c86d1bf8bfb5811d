"""ctypes binding of csrc/libhcm.so (the C ABI of include/hcm.h).  Plumbing only: every call goes straight to the
CUDA library; there is no Python or CPU implementation of the mesher behind it, and importing this module fails
loudly when the extension has not been built."""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from . import registry as _registry

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libhcm.so")

_vp = ctypes.c_void_p


class HcmError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"hcm status {status}: {text}")
        self.status = status


class _Registry(ctypes.Structure):
    _fields_ = [("n_props", ctypes.c_uint32), ("n_meshes", ctypes.c_uint32), ("n_textures", ctypes.c_uint32), ("_pad", ctypes.c_uint32),
                ("prop_index", _vp), ("props", _vp), ("meshes", _vp), ("tex_flags", _vp)]


class _Config(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("origin", ctypes.c_int16 * 3), ("extent", ctypes.c_uint16 * 3),
                ("max_chunks", ctypes.c_uint32), ("max_tasks", ctypes.c_uint32), ("vertex_capacity", ctypes.c_uint64)]


class _BatchView(ctypes.Structure):
    _fields_ = [("n_tasks", ctypes.c_uint32), ("n_meshes", ctypes.c_uint32), ("n_vertices", ctypes.c_uint64), ("n_indices", ctypes.c_uint64),
                ("tasks", _vp), ("meshes", _vp), ("vertices", _vp), ("indices", _vp)]


class _BatchCounts(ctypes.Structure):
    _fields_ = [("n_tasks", ctypes.c_uint32), ("n_meshes", ctypes.c_uint32), ("n_vertices", ctypes.c_uint64), ("n_indices", ctypes.c_uint64),
                ("n_worked", ctypes.c_uint32), ("_pad", ctypes.c_uint32)]


class _EditResult(ctypes.Structure):
    _fields_ = [("n_changed_blocks", ctypes.c_uint32), ("n_sunlight", ctypes.c_uint32), ("n_remesh", ctypes.c_uint32), ("n_pending", ctypes.c_uint32),
                ("sunlight", _vp), ("remesh_pos", _vp), ("remesh_flags", _vp)]


BLOCK_EDIT_DTYPE = np.dtype([("chunk", np.int16, 3), ("index", np.uint16), ("block", np.uint16), ("_pad", np.uint16)])
SUNLIGHT_CHANGE_DTYPE = np.dtype([("chunk", np.int16, 3), ("index", np.uint16), ("old_height", np.int8), ("new_height", np.int8), ("_pad", np.uint16)])
assert BLOCK_EDIT_DTYPE.itemsize == 12 and SUNLIGHT_CHANGE_DTYPE.itemsize == 12


class _Stats(ctypes.Structure):
    _fields_ = [("upload_copy_ms", ctypes.c_float), ("derive_ms", ctypes.c_float), ("classify_ms", ctypes.c_float), ("mesh_ms", ctypes.c_float),
                ("fetch_copy_ms", ctypes.c_float), ("kernel_launches", ctypes.c_uint32), ("h2d_bytes", ctypes.c_uint64),
                ("d2h_bytes", ctypes.c_uint64), ("device_bytes", ctypes.c_uint64), ("sink_ms", ctypes.c_float),
                ("upload_unscanned_chunks", ctypes.c_uint32), ("p2p_bytes", ctypes.c_uint64)]


assert ctypes.sizeof(_Stats) == 64


MAX_VARIANTS = 3


class _KernelTune(ctypes.Structure):
    _fields_ = [("selected", ctypes.c_uint32), ("state", ctypes.c_uint32), ("rounds", ctypes.c_uint32), ("n_variants", ctypes.c_uint32),
                ("baseline_ms", ctypes.c_float), ("variant_ms", ctypes.c_float * MAX_VARIANTS), ("mismatches", ctypes.c_uint32 * MAX_VARIANTS)]


class _TuneInfo(ctypes.Structure):
    _fields_ = [("derive", _KernelTune), ("mesh", _KernelTune)]


assert ctypes.sizeof(_TuneInfo) == 88
TUNE_STATES = ["pending", "variant adopted", "baseline kept", "baseline kept, a variant rejected (bytes differ)", "forced", "off"]
VARIANT_NAMES = {"derive": ["baseline", "quarters + hoist", "hoist + L2 prefetch", "quarters + hoist + L2 prefetch"],
                 "mesh": ["baseline", "merge height test four rows per round", "merge scan two rows per iteration"]}


MESH_DESC_DTYPE = np.dtype([("task", np.uint32), ("transparent", np.uint32), ("vertex_offset", np.uint64), ("index_offset", np.uint64),
                            ("vertex_count", np.uint32), ("index_count", np.uint32), ("aabb_min", np.uint32, 3), ("aabb_max", np.uint32, 3)])
TASK_RESULT_DTYPE = np.dtype([("first_mesh", np.uint32), ("n_meshes", np.uint32)])
assert MESH_DESC_DTYPE.itemsize == 56
MESH_PLACEMENT_DTYPE = np.dtype([("mesh", np.uint32), ("_pad", np.uint32), ("vertex_offset", np.uint64), ("index_offset", np.uint64)])
assert MESH_PLACEMENT_DTYPE.itemsize == 24
HINT_UNKNOWN, HINT_NOT_UNIFORM = -1, -2

EXPORTS = ["hcm_registry_from_blob", "hcm_create", "hcm_destroy", "hcm_last_error", "hcm_status_string", "hcm_host_alloc", "hcm_host_free",
           "hcm_upload_chunks", "hcm_upload_chunks_hinted", "hcm_mesh_counts", "hcm_mesh_emit_into", "hcm_copy_chunks_from", "hcm_remove_chunks", "hcm_rederive_chunks", "hcm_is_resident", "hcm_sync", "hcm_mesh_batch", "hcm_mesh_fetch", "hcm_mesh_wait",
           "hcm_upload_and_mesh", "hcm_set_blocks", "hcm_get_stats", "hcm_stream", "hcm_debug_light", "hcm_get_tune_info", "hcm_set_option"]

_lib = None


def build_library(force: bool = False) -> str:
    """Compile csrc/libhcm.so for sm_100a with nvcc (in-tree; cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", CSRC], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                              "There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.hcm_last_error.restype = ctypes.c_char_p
        L.hcm_last_error.argtypes = [_vp]
        L.hcm_status_string.restype = ctypes.c_char_p
        L.hcm_host_alloc.restype = _vp
        L.hcm_host_alloc.argtypes = [ctypes.c_size_t]
        L.hcm_host_free.argtypes = [_vp]
        L.hcm_stream.restype = _vp
        L.hcm_stream.argtypes = [_vp]
        L.hcm_create.argtypes = [ctypes.POINTER(_Registry), ctypes.POINTER(_Config), ctypes.POINTER(_vp)]
        L.hcm_destroy.argtypes = [_vp]
        L.hcm_registry_from_blob.argtypes = [_vp, ctypes.c_size_t, ctypes.POINTER(_Registry)]
        L.hcm_upload_chunks.argtypes = [_vp, ctypes.c_uint32, _vp, _vp, _vp]
        L.hcm_upload_chunks_hinted.argtypes = [_vp, ctypes.c_uint32, _vp, _vp, _vp, _vp]
        L.hcm_mesh_counts.argtypes = [_vp, ctypes.POINTER(_BatchView)]
        L.hcm_mesh_emit_into.argtypes = [_vp, ctypes.c_uint32, _vp, _vp, ctypes.c_uint64, _vp, ctypes.c_uint64]
        L.hcm_remove_chunks.argtypes = [_vp, ctypes.c_uint32, _vp]
        L.hcm_copy_chunks_from.argtypes = [_vp, _vp, ctypes.c_uint32, _vp]
        L.hcm_rederive_chunks.argtypes = [_vp, ctypes.c_uint32, _vp]
        L.hcm_is_resident.argtypes = [_vp, _vp]
        L.hcm_sync.argtypes = [_vp]
        L.hcm_mesh_batch.argtypes = [_vp, ctypes.c_uint32, _vp]
        L.hcm_mesh_fetch.argtypes = [_vp, ctypes.POINTER(_BatchView)]
        L.hcm_mesh_wait.argtypes = [_vp, ctypes.POINTER(_BatchCounts)]
        L.hcm_upload_and_mesh.argtypes = [_vp, ctypes.c_uint32, _vp, _vp, _vp, ctypes.c_uint32, _vp, ctypes.POINTER(_BatchView)]
        L.hcm_set_blocks.argtypes = [_vp, ctypes.c_uint32, _vp, ctypes.POINTER(_EditResult)]
        L.hcm_get_stats.argtypes = [_vp, ctypes.POINTER(_Stats)]
        L.hcm_debug_light.argtypes = [_vp, _vp, _vp]
        L.hcm_get_tune_info.argtypes = [_vp, ctypes.POINTER(_TuneInfo)]
        L.hcm_set_option.argtypes = [_vp, ctypes.c_char_p, ctypes.c_int64]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


class PinnedArray:
    """numpy view over pinned host memory from hcm_host_alloc (truly asynchronous upload source)."""

    def __init__(self, shape, dtype):
        L = load_library()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = L.hcm_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError("hcm_host_alloc failed")
        buf = (ctypes.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            load_library().hcm_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class BatchResult:
    """Host views of one fetched batch (valid until the next batch on the same Mesher)."""
    tasks: np.ndarray  # TASK_RESULT_DTYPE[n_tasks]
    meshes: np.ndarray  # MESH_DESC_DTYPE[n_meshes]
    vertices: np.ndarray  # u64[n_vertices]
    indices: np.ndarray  # u16[n_indices]

    def task_meshes(self, t: int) -> list[dict]:
        """The std::vector<BlockMesh> of task t, as PushChunkMesh would receive it."""
        first, n = int(self.tasks[t]["first_mesh"]), int(self.tasks[t]["n_meshes"])
        out = []
        for d in self.meshes[first:first + n]:
            vo, io = int(d["vertex_offset"]), int(d["index_offset"])
            out.append(dict(transparent=int(d["transparent"]), aabb_min=d["aabb_min"].copy(), aabb_max=d["aabb_max"].copy(),
                            vertices=self.vertices[vo:vo + int(d["vertex_count"])], indices=self.indices[io:io + int(d["index_count"])]))
        return out

    def serialise(self, t: int) -> np.ndarray:
        """Task t in the oracle's serialised mesh-list layout (oracle/ref_shim.cpp) for byte-for-byte comparison."""
        ms = self.task_meshes(t)
        parts = [np.array([len(ms)], dtype=np.uint32).view(np.uint8)]
        for m in ms:
            hdr = np.concatenate([[m["transparent"]], m["aabb_min"], m["aabb_max"], [len(m["vertices"]), len(m["indices"])]]).astype(np.uint32)
            parts.append(hdr.view(np.uint8))
        for m in ms:
            parts.append(np.ascontiguousarray(m["vertices"]).view(np.uint8))
            ib = np.ascontiguousarray(m["indices"]).view(np.uint8)
            parts.append(ib)
            if ib.size & 3:
                parts.append(np.zeros(4 - (ib.size & 3), dtype=np.uint8))
        return np.concatenate(parts)


class Mesher:
    """One hcm_ctx: a device-resident world window on one GPU plus its batch output arenas."""

    def __init__(self, origin, extent, device: int = 0, max_chunks: int = 0, max_tasks: int = 0, vertex_capacity: int = 0,
                 registry_blob: np.ndarray | None = None):
        self.L = load_library()
        self.blob = np.ascontiguousarray(registry_blob if registry_blob is not None else _registry.load().blob, dtype=np.uint8)
        reg = _Registry()
        rc = self.L.hcm_registry_from_blob(_ptr(self.blob), self.blob.size, ctypes.byref(reg))
        if rc:
            raise HcmError(rc, "bad registry blob")
        cfg = _Config()
        cfg.device = device
        for i in range(3):
            cfg.origin[i] = int(origin[i])
            cfg.extent[i] = int(extent[i])
        cfg.max_chunks, cfg.max_tasks, cfg.vertex_capacity = int(max_chunks), int(max_tasks), int(vertex_capacity)
        ctx = _vp()
        rc = self.L.hcm_create(ctypes.byref(reg), ctypes.byref(cfg), ctypes.byref(ctx))
        if rc:
            raise HcmError(rc, (self.L.hcm_last_error(None) or b"").decode())
        self.ctx = ctx
        self._keep = []
        self.origin = tuple(int(v) for v in origin)
        self.extent = tuple(int(v) for v in extent)

    def close(self):
        if getattr(self, "ctx", None):
            self.L.hcm_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise HcmError(rc, (self.L.hcm_last_error(self.ctx) or b"").decode())

    # ---- residency -------------------------------------------------------------------------------------------------
    def upload(self, positions, blocks, heights, uniform_hint=None) -> None:
        """hcm_upload_chunks; with ``uniform_hint`` (int32 per chunk: a Block value, HINT_UNKNOWN or HINT_NOT_UNIFORM)
        hcm_upload_chunks_hinted."""
        pos = np.ascontiguousarray(positions, dtype=np.int16).reshape(-1, 3)
        b = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(len(pos), 32768)
        h = np.ascontiguousarray(heights, dtype=np.int8).reshape(len(pos), 1024)
        # the C side reads the arrays asynchronously until the next sync / fetch: keep every one of them until then
        self._keep.append((pos, b, h))
        if uniform_hint is None:
            self._check(self.L.hcm_upload_chunks(self.ctx, len(pos), _ptr(pos), _ptr(b), _ptr(h)))
        else:
            hint = np.ascontiguousarray(uniform_hint, dtype=np.int32).reshape(len(pos))
            self._check(self.L.hcm_upload_chunks_hinted(self.ctx, len(pos), _ptr(pos), _ptr(b), _ptr(h), _ptr(hint)))

    def copy_chunks_from(self, src: "Mesher", positions) -> None:
        """hcm_copy_chunks_from: chunks resident in ``src`` become resident here, device to device (halo replication between GPUs)."""
        pos = np.ascontiguousarray(positions, dtype=np.int16).reshape(-1, 3)
        self._check(self.L.hcm_copy_chunks_from(self.ctx, src.ctx, len(pos), _ptr(pos)))

    def remove(self, positions) -> None:
        pos = np.ascontiguousarray(positions, dtype=np.int16).reshape(-1, 3)
        self._check(self.L.hcm_remove_chunks(self.ctx, len(pos), _ptr(pos)))

    def rederive(self, positions=None) -> None:
        """Re-run the derivation kernel from device-resident Block data (None: every resident chunk)."""
        if positions is None:
            self._check(self.L.hcm_rederive_chunks(self.ctx, 0, None))
        else:
            pos = np.ascontiguousarray(positions, dtype=np.int16).reshape(-1, 3)
            self._check(self.L.hcm_rederive_chunks(self.ctx, len(pos), _ptr(pos)))

    def set_blocks(self, edit_list) -> dict:
        """Block edits on the resident world (hcm_set_blocks).  ``edit_list`` = [((cx,cy,cz), index, block), ...] or a
        BLOCK_EDIT_DTYPE array.  Returns n_changed_blocks, n_pending, ``sunlight`` (SUNLIGHT_CHANGE_DTYPE copies),
        ``remesh`` (int16[n,3]: the mesh tasks to submit) and ``remesh_flags`` (u8[n]: 1 = pushed by a block edit -- high priority in
        the reference --, 2 = pushed by a sunlight change)."""
        if isinstance(edit_list, np.ndarray) and edit_list.dtype == BLOCK_EDIT_DTYPE:
            e = np.ascontiguousarray(edit_list)
        else:
            e = np.zeros(len(edit_list), dtype=BLOCK_EDIT_DTYPE)
            for i, (cpos, index, block) in enumerate(edit_list):
                e[i]["chunk"], e[i]["index"], e[i]["block"] = cpos, index, block
        r = _EditResult()
        self._check(self.L.hcm_set_blocks(self.ctx, len(e), _ptr(e) if len(e) else None, ctypes.byref(r)))

        def copy(ptr, count, dtype):
            if not count:
                return np.zeros(0, dtype=dtype)
            buf = (ctypes.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype, count=count).copy()

        return dict(n_changed_blocks=r.n_changed_blocks, n_pending=r.n_pending, sunlight=copy(r.sunlight, r.n_sunlight, SUNLIGHT_CHANGE_DTYPE),
                    remesh=copy(r.remesh_pos, r.n_remesh * 3, np.int16).reshape(-1, 3), remesh_flags=copy(r.remesh_flags, r.n_remesh, np.uint8))

    def is_resident(self, pos) -> bool:
        p = np.ascontiguousarray(pos, dtype=np.int16)
        return bool(self.L.hcm_is_resident(self.ctx, _ptr(p)))

    def sync(self) -> None:
        self._check(self.L.hcm_sync(self.ctx))
        self._keep.clear()

    # ---- meshing ---------------------------------------------------------------------------------------------------
    def mesh_async(self, positions) -> None:
        pos = np.ascontiguousarray(positions, dtype=np.int16).reshape(-1, 3)
        self._check(self.L.hcm_mesh_batch(self.ctx, len(pos), _ptr(pos)))

    def wait(self) -> dict:
        c = _BatchCounts()
        self._check(self.L.hcm_mesh_wait(self.ctx, ctypes.byref(c)))
        self._keep.clear()
        return dict(n_tasks=c.n_tasks, n_meshes=c.n_meshes, n_vertices=c.n_vertices, n_indices=c.n_indices, n_worked=c.n_worked)

    @staticmethod
    def _result(v: _BatchView) -> BatchResult:
        def view(ptr, count, dtype):
            if not count or not ptr:  # hcm_mesh_counts returns totals without arrays
                return np.zeros(0, dtype=dtype)
            nbytes = count * np.dtype(dtype).itemsize
            buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype, count=count)

        return BatchResult(view(v.tasks, v.n_tasks, TASK_RESULT_DTYPE), view(v.meshes, v.n_meshes, MESH_DESC_DTYPE),
                           view(v.vertices, v.n_vertices, np.uint64), view(v.indices, v.n_indices, np.uint16))

    def fetch(self) -> BatchResult:
        v = _BatchView()
        self._check(self.L.hcm_mesh_fetch(self.ctx, ctypes.byref(v)))
        self._keep.clear()
        return self._result(v)

    # ---- device-side sink (hcm_mesh_counts + hcm_mesh_emit_into) ---------------------------------------------------
    def counts(self) -> BatchResult:
        """Tasks and mesh descriptors of the last batch; no vertex or index byte is copied (``vertices`` / ``indices`` are empty)."""
        v = _BatchView()
        self._check(self.L.hcm_mesh_counts(self.ctx, ctypes.byref(v)))
        self._keep.clear()
        r = self._result(v)
        self.n_vertices, self.n_indices = int(v.n_vertices), int(v.n_indices)
        return BatchResult(r.tasks, r.meshes, np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint16))

    def emit_into(self, placements, dev_vertex_ptr: int, vertex_capacity: int, dev_index_ptr: int, index_capacity: int) -> None:
        """Copy the placed meshes into caller-owned DEVICE buffers (raw device pointers, capacities in vertices / indices)."""
        pl = np.ascontiguousarray(placements, dtype=MESH_PLACEMENT_DTYPE)
        self._check(self.L.hcm_mesh_emit_into(self.ctx, len(pl), _ptr(pl) if len(pl) else None, _vp(dev_vertex_ptr), int(vertex_capacity),
                                              _vp(dev_index_ptr), int(index_capacity)))

    def upload_and_mesh(self, positions, blocks, heights, task_positions) -> BatchResult:
        """upload() + mesh_async() + fetch() with host->device copies, kernels and device->host copies overlapped
        (hcm_upload_and_mesh).  Result rows follow ``task_positions``."""
        pos = np.ascontiguousarray(positions, dtype=np.int16).reshape(-1, 3)
        b = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(len(pos), 32768)
        h = np.ascontiguousarray(heights, dtype=np.int8).reshape(len(pos), 1024)
        tp = np.ascontiguousarray(task_positions, dtype=np.int16).reshape(-1, 3)
        v = _BatchView()
        self._check(self.L.hcm_upload_and_mesh(self.ctx, len(pos), _ptr(pos), _ptr(b), _ptr(h), len(tp), _ptr(tp), ctypes.byref(v)))
        return self._result(v)

    def mesh(self, positions) -> BatchResult:
        self.mesh_async(positions)
        return self.fetch()

    # ---- measurement / debugging -----------------------------------------------------------------------------------
    def stats(self) -> dict:
        s = _Stats()
        self._check(self.L.hcm_get_stats(self.ctx, ctypes.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in _Stats._fields_}

    def tune_info(self) -> dict:
        """hcm_get_tune_info: which form of derive_kernel / mesh_kernel this context launches and what the A/B rounds measured."""
        t = _TuneInfo()
        self._check(self.L.hcm_get_tune_info(self.ctx, ctypes.byref(t)))
        out = {}
        for name in ("derive", "mesh"):
            k = getattr(t, name)
            n = int(k.n_variants)
            out[name] = {"selected": int(k.selected), "selected_name": VARIANT_NAMES[name][k.selected] if k.selected < len(VARIANT_NAMES[name]) else str(k.selected),
                         "state": TUNE_STATES[k.state] if k.state < len(TUNE_STATES) else str(k.state), "rounds": int(k.rounds), "n_variants": n,
                         "baseline_ms": float(k.baseline_ms), "variant_ms": [float(k.variant_ms[v]) for v in range(n)],
                         "mismatches": [int(k.mismatches[v]) for v in range(n)],
                         "variant_names": VARIANT_NAMES[name][1:n + 1], "times": "summed over the A/B rounds (CUDA events)"}
        return out

    def set_option(self, name: str, value: int) -> None:
        self._check(self.L.hcm_set_option(self.ctx, name.encode(), int(value)))

    def stream_ptr(self) -> int:
        return int(self.L.hcm_stream(self.ctx) or 0)

    def debug_light(self, pos) -> np.ndarray:
        p = np.ascontiguousarray(pos, dtype=np.int16)
        out = np.zeros(34 * 34 * 34, dtype=np.uint8)
        self._check(self.L.hcm_debug_light(self.ctx, _ptr(p), _ptr(out)))
        return out
