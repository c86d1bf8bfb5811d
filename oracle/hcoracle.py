"""TEST INFRASTRUCTURE -- ctypes wrappers over the parity oracles.  Not part of the shipped product.

* ``RefWorld``  : oracle/_ref/libhcref.so  -- the UNMODIFIED reference mesher / terrain / registry
                  (built by ``make -C oracle ref`` from /root/reference; the prebuilt .so travels to the GPU box).
* ``PortOracle``: oracle/libhcoracle.so    -- this repo's CPU restatement (oracle/hc_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py (input generation, cpu_baseline, --impl reference) import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libhcref.so")
PORT_SO = os.path.join(HERE, "libhcoracle.so")

_vp = ctypes.c_void_p


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def build(ref: bool = True, port: bool = True) -> None:
    """Compile the oracles (reference only when /root/reference is present)."""
    targets = []
    if port:
        targets.append("port")
    if ref:
        targets.append("ref")
    subprocess.run(["make", "-C", HERE, "-j8"] + targets, check=True, stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def parse_meshes(buf: np.ndarray) -> list[dict]:
    """Decode the serialised mesh list (layout documented in oracle/ref_shim.cpp)."""
    n = int(buf[:4].view(np.uint32)[0])
    hdr = buf[4:4 + 36 * n].view(np.uint32).reshape(n, 9)
    off = 4 + 36 * n
    out = []
    for i in range(n):
        nv, ni = int(hdr[i, 7]), int(hdr[i, 8])
        v = buf[off:off + nv * 8].view(np.uint64).copy()
        off += nv * 8
        idx = buf[off:off + ni * 2].view(np.uint16).copy()
        off += (ni * 2 + 3) & ~3
        out.append(dict(transparent=int(hdr[i, 0]), aabb_min=hdr[i, 1:4].copy(), aabb_max=hdr[i, 4:7].copy(),
                        vertices=v, indices=idx))
    return out


class RefWorld:
    """A set of reference ``hc::client::Chunk`` objects + the reference mesher over them."""

    def __init__(self):
        if not have_ref():
            raise RuntimeError("oracle/_ref/libhcref.so missing: run `make -C oracle ref` where /root/reference exists")
        L = ctypes.CDLL(REF_SO)
        L.hcref_world_create.restype = _vp
        L.hcref_world_chunk_count.restype = ctypes.c_uint64
        L.hcref_mesh.restype = ctypes.c_int64
        L.hcref_result_copy.restype = ctypes.c_int64
        L.hcref_registry_blob.restype = ctypes.c_int64
        self.L = L
        self.w = _vp(L.hcref_world_create())

    def __del__(self):
        try:
            self.L.hcref_world_destroy(self.w)
        except Exception:
            pass

    @staticmethod
    def hardware_threads() -> int:
        return os.cpu_count() or 1

    def registry_blob(self) -> np.ndarray:
        need = self.L.hcref_registry_blob(None, ctypes.c_uint64(0))
        if need < 0:
            raise RuntimeError(f"registry export failed: {need}")
        blob = np.zeros(need, dtype=np.uint8)
        self.L.hcref_registry_blob(_ptr(blob), ctypes.c_uint64(need))
        return blob

    def generate(self, positions: np.ndarray, seed: int = 12314524, threads: int = 0) -> None:
        pos = np.ascontiguousarray(positions, dtype=np.int16).reshape(-1, 3)
        self.L.hcref_world_generate(self.w, ctypes.c_uint32(seed), _ptr(pos), ctypes.c_uint64(len(pos)),
                                    ctypes.c_int(threads or self.hardware_threads()))

    def set_chunk(self, pos, blocks: np.ndarray, heights: np.ndarray) -> None:
        p = np.ascontiguousarray(pos, dtype=np.int16)
        b = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(-1)
        h = np.ascontiguousarray(heights, dtype=np.int8).reshape(-1)
        assert b.size == 32768 and h.size == 1024
        self.L.hcref_world_set_chunk(self.w, _ptr(p), _ptr(b), _ptr(h))

    def get_chunk(self, pos):
        p = np.ascontiguousarray(pos, dtype=np.int16)
        b = np.empty(32768, dtype=np.uint16)
        h = np.empty(1024, dtype=np.int8)
        if self.L.hcref_world_get_chunk(self.w, _ptr(p), _ptr(b), _ptr(h)) != 0:
            raise KeyError(tuple(int(x) for x in p))
        return b, h

    def get_chunks(self, positions: np.ndarray, blocks_out: np.ndarray | None = None, heights_out: np.ndarray | None = None):
        pos = np.ascontiguousarray(positions, dtype=np.int16).reshape(-1, 3)
        b = blocks_out if blocks_out is not None else np.empty((len(pos), 32768), dtype=np.uint16)
        h = heights_out if heights_out is not None else np.empty((len(pos), 1024), dtype=np.int8)
        assert b.flags.c_contiguous and h.flags.c_contiguous and b.dtype == np.uint16 and h.dtype == np.int8
        self.L.hcref_world_get_chunks.restype = ctypes.c_int64
        missing = self.L.hcref_world_get_chunks(self.w, _ptr(pos), ctypes.c_uint64(len(pos)), _ptr(b), _ptr(h))
        if missing:
            raise KeyError(f"{missing} chunks missing")
        return b, h

    def mesh(self, pos) -> np.ndarray:
        """Unmodified ``ChunkTaskRunner<kMesh>::Run`` on one chunk; returns the serialised meshes."""
        p = np.ascontiguousarray(pos, dtype=np.int16)
        n = self.L.hcref_mesh(self.w, _ptr(p))
        if n < 0:
            raise KeyError("a neighbour chunk is missing")
        buf = np.zeros(n, dtype=np.uint8)
        self.L.hcref_result_copy(self.w, _ptr(buf))
        return buf

    def mesh_batch(self, positions: np.ndarray, threads: int = 0, per_chunk: bool = False):
        """Multithreaded reference mesher (one runner per thread). Returns (stats dict, per-chunk u64[n,4] or None)."""

        class Stats(ctypes.Structure):
            _fields_ = [("seconds", ctypes.c_double), ("chunks", ctypes.c_uint64), ("meshes", ctypes.c_uint64),
                        ("vertices", ctypes.c_uint64), ("indices", ctypes.c_uint64), ("threads", ctypes.c_int32),
                        ("missing", ctypes.c_int32)]

        pos = np.ascontiguousarray(positions, dtype=np.int16).reshape(-1, 3)
        pc = np.zeros((len(pos), 4), dtype=np.uint64) if per_chunk else None
        st = Stats()
        self.L.hcref_mesh_batch(self.w, _ptr(pos), ctypes.c_uint64(len(pos)), ctypes.c_int(threads or self.hardware_threads()),
                                _ptr(pc) if per_chunk else None, ctypes.byref(st))
        return {f[0]: getattr(st, f[0]) for f in Stats._fields_}, pc

    def light34(self, pos) -> np.ndarray:
        p = np.ascontiguousarray(pos, dtype=np.int16)
        out = np.zeros(34 * 34 * 34, dtype=np.uint8)
        if self.L.hcref_light34(self.w, _ptr(p), _ptr(out)) != 0:
            raise KeyError("a neighbour chunk is missing")
        return out

    def dynamic_mesh_eval(self, block: int, neighbours: np.ndarray):
        nb = np.ascontiguousarray(neighbours, dtype=np.uint16)
        faces = np.zeros(16 * 24, dtype=np.uint8)
        tex = np.zeros(6, dtype=np.uint16)
        dyn = ctypes.c_int(0)
        n = self.L.hcref_dynamic_mesh_eval(ctypes.c_uint16(block), _ptr(nb), _ptr(faces), _ptr(tex), ctypes.byref(dyn))
        return n, faces[:max(n, 0) * 24].copy(), tex, dyn.value


class RefEditWorld:
    """The reference's own block-edit chain: a reference World holding the caller's chunks, edits pushed as World::SetBlockBulk does,
    the reference's ChunkTaskPool::Run run to quiescence on one thread (oracle/ref_shim.cpp: hcref_edit_*).  Pins hcm_set_blocks and
    oracle/edit_oracle.py to compiled reference code (kSetBlock, kFloodSunlight, kSetSunlight and the kMesh tasks they push)."""

    EDIT_DTYPE = np.dtype([("chunk", np.int16, 3), ("index", np.uint16), ("block", np.uint16), ("_pad", np.uint16)])

    def __init__(self, load_radius: int = 4):
        if not have_ref():
            raise RuntimeError("oracle/_ref/libhcref.so missing: run `make -C oracle ref` where /root/reference exists")
        L = ctypes.CDLL(REF_SO)
        L.hcref_edit_create.restype = _vp
        L.hcref_edit_apply.restype = ctypes.c_int64
        L.hcref_edit_pushed_mesh.restype = ctypes.c_int64
        self.L = L
        self.w = _vp(L.hcref_edit_create(ctypes.c_int(load_radius)))

    def __del__(self):
        try:
            self.L.hcref_edit_destroy(self.w)
        except Exception:
            pass

    def set_chunk(self, pos, blocks, heights) -> None:
        p = np.ascontiguousarray(pos, dtype=np.int16)
        b = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(-1)
        h = np.ascontiguousarray(heights, dtype=np.int8).reshape(-1)
        if self.L.hcref_edit_set_chunk(self.w, _ptr(p), _ptr(b), _ptr(h)) != 0:
            raise KeyError(f"{tuple(int(v) for v in p)} is outside the reference world's load sphere")

    def get_chunk(self, pos):
        p = np.ascontiguousarray(pos, dtype=np.int16)
        b = np.empty(32768, dtype=np.uint16)
        h = np.empty(1024, dtype=np.int8)
        if self.L.hcref_edit_get_chunk(self.w, _ptr(p), _ptr(b), _ptr(h)) != 0:
            raise KeyError(tuple(int(x) for x in p))
        return b, h

    def apply(self, edit_list, with_meshes: bool = False):
        """edit_list = [((cx,cy,cz), index, block), ...].  -> (pushed int16[n,3] in the order the mesh tasks ran, high u8[n],
        n_still_queued, [serialised meshes per pushed task] or None)."""
        e = np.zeros(len(edit_list), dtype=self.EDIT_DTYPE)
        for i, (cpos, index, block) in enumerate(edit_list):
            e[i]["chunk"], e[i]["index"], e[i]["block"] = cpos, index, block
        cap = 4096
        pos = np.zeros((cap, 3), dtype=np.int16)
        high = np.zeros(cap, dtype=np.uint8)
        still = ctypes.c_uint64(0)
        n = self.L.hcref_edit_apply(self.w, _ptr(e), ctypes.c_uint64(len(e)), _ptr(pos), _ptr(high), ctypes.c_uint64(cap), ctypes.byref(still))
        assert 0 <= n <= cap
        meshes = None
        if with_meshes:
            meshes = []
            for i in range(n):
                size = self.L.hcref_edit_pushed_mesh(self.w, ctypes.c_uint64(i), None)
                buf = np.zeros(size, dtype=np.uint8)
                self.L.hcref_edit_pushed_mesh(self.w, ctypes.c_uint64(i), _ptr(buf))
                meshes.append(buf)
        return pos[:n].copy(), high[:n].copy(), int(still.value), meshes


BIND_SO = os.path.join(HERE, "_ref", "libhcbind.so")


def have_binding() -> bool:
    return os.path.exists(BIND_SO)


class BindingWorld:
    """oracle/_ref/libhcbind.so: the UNMODIFIED reference client (task pool, chunk pool, world, edit tasks) with client/src/ChunkMeshTask.cpp
    replaced by hypercraft_b200/host/reference_binding/ChunkMeshTask.cpp -- the drop-in mesh lane on top of the C ABI -- compiled
    against the reference's own Chunk / ChunkTaskPool / WorldRenderer types (make -C oracle binding).  `hcm_library` is the build of
    the C ABI the binding's hcm_* calls resolve to (libhcm.so on a GPU box; the CPU emulation build in the CPU test tier)."""

    def __init__(self, hcm_library: str, origin, extent, load_radius: int, device: int = 0):
        if not have_binding():
            raise RuntimeError("oracle/_ref/libhcbind.so missing: run `make -C oracle binding` where /root/reference exists")
        self._hcm = ctypes.CDLL(hcm_library, mode=ctypes.RTLD_GLOBAL)
        L = ctypes.CDLL(BIND_SO)
        L.hcref_edit_create.restype = _vp
        L.hcref_edit_mesh.restype = ctypes.c_int64
        L.hcref_edit_pushed_mesh.restype = ctypes.c_int64
        self.L = L
        o = np.ascontiguousarray(origin, dtype=np.int16)
        e = np.ascontiguousarray(extent, dtype=np.uint16)
        err = ctypes.create_string_buffer(512)
        rc = L.hcref_binding_configure(ctypes.c_int(device), _ptr(o), _ptr(e), err, ctypes.c_uint64(512))
        if rc:
            raise RuntimeError(f"hcm_binding::Configure failed with status {rc}: {err.value.decode()}")
        self.w = _vp(L.hcref_edit_create(ctypes.c_int(load_radius)))

    def close(self):
        if getattr(self, "w", None):
            self.L.hcref_edit_destroy(self.w)
            self.L.hcref_binding_shutdown()
            self.w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_chunk(self, pos, blocks, heights) -> None:
        p = np.ascontiguousarray(pos, dtype=np.int16)
        b = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(-1)
        h = np.ascontiguousarray(heights, dtype=np.int8).reshape(-1)
        if self.L.hcref_edit_set_chunk(self.w, _ptr(p), _ptr(b), _ptr(h)) != 0:
            raise KeyError(f"{tuple(int(v) for v in p)} is outside the reference world's load sphere")

    def mesh(self, positions, high: bool = False):
        """Push<kMesh> for every position, the reference's ChunkTaskPool::Run to quiescence.  -> {position: serialised meshes} of the
        PushChunkMesh calls the (shadow) WorldRenderer received, and the number of calls."""
        pos = np.ascontiguousarray(positions, dtype=np.int16).reshape(-1, 3)
        cap = 4 * len(pos) + 64
        out = np.zeros((cap, 3), dtype=np.int16)
        n = self.L.hcref_edit_mesh(self.w, _ptr(pos), ctypes.c_uint64(len(pos)), ctypes.c_int(1 if high else 0), _ptr(out), ctypes.c_uint64(cap))
        assert 0 <= n <= cap
        got = {}
        for i in range(n):
            size = self.L.hcref_edit_pushed_mesh(self.w, ctypes.c_uint64(i), None)
            buf = np.zeros(size, dtype=np.uint8)
            self.L.hcref_edit_pushed_mesh(self.w, ctypes.c_uint64(i), _ptr(buf))
            got[tuple(int(v) for v in out[i])] = buf
        return got, int(n)


_port_lib = None


def _port():
    global _port_lib
    if _port_lib is None:
        if not os.path.exists(PORT_SO):
            build(ref=False, port=True)
        L = ctypes.CDLL(PORT_SO)
        L.hco_create.restype = _vp
        L.hco_mesh.restype = ctypes.c_int64
        L.hco_fnv1a.restype = ctypes.c_uint64
        L.hco_fnv1a.argtypes = [_vp, ctypes.c_uint64]
        L.hco_digest_batch_view.argtypes = [_vp, ctypes.c_uint64, _vp, _vp, _vp, _vp, ctypes.c_int]
        _port_lib = L
    return _port_lib


def fnv1a(buf: np.ndarray) -> int:
    """FNV-1a 64 of a byte array: the per-chunk digest RefWorld.mesh_batch(per_chunk=True) returns in column 3."""
    b = np.ascontiguousarray(buf, dtype=np.uint8)
    return int(_port().hco_fnv1a(_ptr(b), ctypes.c_uint64(b.size)))


def digest_batch(result, threads: int = 0) -> np.ndarray:
    """Checker for whole batches.  ``result`` = a fetched batch (hypercraft_b200.mesher.BatchResult: tasks, meshes, vertices,
    indices as the C ABI returned them) -> u64[n_tasks, 4] = {n_meshes, n_vertices, n_indices, FNV-1a of the task's serialised
    meshes}: row for row what RefWorld.mesh_batch(per_chunk=True) / PortOracle.mesh_batch give for the same chunks."""
    n = len(result.tasks)
    out = np.zeros((n, 4), dtype=np.uint64)
    if n:
        t, m = np.ascontiguousarray(result.tasks), np.ascontiguousarray(result.meshes)
        v, i = np.ascontiguousarray(result.vertices), np.ascontiguousarray(result.indices)
        _port().hco_digest_batch_view(_ptr(t), ctypes.c_uint64(n), _ptr(m), _ptr(v), _ptr(i), _ptr(out), ctypes.c_int(threads or (os.cpu_count() or 1)))
    return out


# neighbour order of the 27-array: NeighbourIndex2CmpXYZ, common/include/common/Position.hpp:146-152
def neighbour_offsets() -> np.ndarray:
    rev = (1, -1, 0)
    return np.array([[rev[i // 9], rev[(i // 3) % 3], rev[i % 3]] for i in range(27)], dtype=np.int16)


class PortOracle:
    """This repo's CPU restatement of the mesh task (oracle/hc_oracle.cpp)."""

    def __init__(self, registry_blob: np.ndarray):
        L = _port()
        self.L = L
        self.blob = np.ascontiguousarray(registry_blob, dtype=np.uint8)
        self.o = _vp(L.hco_create(_ptr(self.blob), ctypes.c_uint64(self.blob.size)))
        if not self.o:
            raise RuntimeError("bad registry blob")
        self.offsets = neighbour_offsets()

    def __del__(self):
        try:
            self.L.hco_destroy(self.o)
        except Exception:
            pass

    def _arrays(self, chunks: dict, pos):
        """chunks: {(x,y,z): (blocks u16[32768], heights i8[1024])}"""
        bp = (_vp * 27)()
        hp = (_vp * 27)()
        keep = []
        for i in range(27):
            key = tuple(int(pos[k]) + int(self.offsets[i, k]) for k in range(3))
            b, h = chunks[key]
            b = np.ascontiguousarray(b, dtype=np.uint16).reshape(-1)
            h = np.ascontiguousarray(h, dtype=np.int8).reshape(-1)
            keep.append((b, h))
            bp[i] = b.ctypes.data
            hp[i] = h.ctypes.data
        return bp, hp, keep

    def mesh(self, chunks: dict, pos) -> np.ndarray:
        bp, hp, keep = self._arrays(chunks, pos)
        n = self.L.hco_mesh(self.o, bp, hp, None, ctypes.c_uint64(0))
        buf = np.zeros(n, dtype=np.uint8)
        self.L.hco_mesh(self.o, bp, hp, _ptr(buf), ctypes.c_uint64(n))
        return buf

    def light34(self, chunks: dict, pos) -> np.ndarray:
        bp, hp, keep = self._arrays(chunks, pos)
        out = np.zeros(34 * 34 * 34, dtype=np.uint8)
        self.L.hco_light34(self.o, bp, hp, _ptr(out))
        return out

    def mesh_batch(self, chunks: dict, positions, threads: int = 1, keep: bool = False):
        """Mesh many chunks on `threads` host threads (hco_mesh holds no state between calls and ctypes releases the GIL).
        -> ({"chunks", "seconds", "vertices"}, per-chunk [n_meshes, vertices, indices, FNV-1a of the serialised meshes] u64,
        serialised buffers or None).
        The timed region holds the mesher calls only, like RefWorld.mesh_batch."""
        import time
        from concurrent.futures import ThreadPoolExecutor

        positions = [tuple(int(c) for c in p) for p in positions]
        args = [self._arrays(chunks, p) for p in positions]
        per = np.zeros((len(positions), 4), dtype=np.uint64)
        bufs = [None] * len(positions)

        def work(first):
            buf = np.zeros(1 << 22, dtype=np.uint8)
            for i in range(first, len(positions), threads):  # interleaved: heavy chunks are neighbours in the task list
                bp, hp, _ = args[i]
                n = self.L.hco_mesh(self.o, bp, hp, _ptr(buf), ctypes.c_uint64(buf.size))
                if n > buf.size:
                    buf = np.zeros(int(n), dtype=np.uint8)
                    self.L.hco_mesh(self.o, bp, hp, _ptr(buf), ctypes.c_uint64(buf.size))
                k = int(buf[:4].view(np.uint32)[0])
                hdr = buf[4:4 + 36 * k].view(np.uint32).reshape(k, 9)
                per[i] = (k, int(hdr[:, 7].sum()), int(hdr[:, 8].sum()), int(self.L.hco_fnv1a(_ptr(buf), ctypes.c_uint64(int(n)))))
                if keep:
                    bufs[i] = buf[:n].copy()

        threads = max(1, min(threads, len(positions)))
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, range(threads)))
        secs = time.perf_counter() - t0
        return {"chunks": len(positions), "seconds": secs, "vertices": int(per[:, 1].sum())}, per, (bufs if keep else None)

    def dynamic_mesh_eval(self, block: int, neighbours: np.ndarray):
        nb = np.ascontiguousarray(neighbours, dtype=np.uint16)
        faces = np.zeros(16 * 24, dtype=np.uint8)
        tex = np.zeros(6, dtype=np.uint16)
        dyn = ctypes.c_int(0)
        n = self.L.hco_dynamic_mesh_eval(self.o, ctypes.c_uint16(block), _ptr(nb), _ptr(faces), _ptr(tex), ctypes.byref(dyn))
        return n, faces[:max(n, 0) * 24].copy(), tex, dyn.value
