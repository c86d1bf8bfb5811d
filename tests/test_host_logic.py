"""Host-side logic and the C-ABI surface (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_load_radius_counts_match_reference_list():
    """client/src/WorldLoadingList.inl end table: r=11 -> 8121 resident, r=20 -> 41443."""
    from hypercraft_b200 import worldgen

    for r, meshed, resident in [(0, 1, 27), (11, 5575, 8121), (20, 33401, 41443)]:
        m, res = worldgen.load_radius_positions(r)
        assert (len(m), len(res)) == (meshed, resident)
        s = {tuple(p) for p in res.tolist()}
        for p in m[:: max(1, len(m) // 50)]:
            assert all(tuple(q) in s for q in worldgen.neighbourhood(p).tolist())


def test_make_block_follows_block_ctor():
    from hypercraft_b200 import worldgen as w

    assert w.make_block(w.STONE) == 1
    assert w.make_block(w.LOG, 2, 1) == 17 | ((2 | (1 << 3)) << 8)
    assert w.make_block(w.FLOWING_WATER, 8) == 15 | (8 << 8)
    assert w.make_block(w.VINE, 0, 5) == 21 | (5 << 8)


def test_registry_blob_is_consistent():
    from hypercraft_b200 import registry

    reg = registry.load()
    assert len(reg.props) == 67 and len(reg.meshes) == 24 and len(reg.tex_flags) == 53
    air = reg.props[reg.prop_index[0]]
    assert air["transparency"] == 2 and air["light_level"] == 0 and not air["tex"].any() and air["liquid_obstacle"] == 0  # block/test/test_block.cpp:10-23
    glow = reg.props[reg.prop_index[24]]
    assert glow["light_level"] == 15 and glow["transparency"] == 0
    # vertical-pass => indirect-pass for every property (block/test/test_block.cpp:58-67)
    assert ((reg.props["transparency"] == 2) <= (reg.props["transparency"] != 0)).all()
    flow = reg.meshes[reg.props[reg.prop_index[15 | (3 << 8)]]["custom_mesh"] - 1]
    assert flow["dynamic_kind"] == 1 and flow["liquid_variant"] == 3 and flow["liquid_height_unit"] == 2
    # meta above the table clamps to the last variant, raw meta is still kept by the ABI (Block.cpp.in:9-16)
    assert reg.prop_index[15 | (200 << 8)] == reg.prop_index[15 | (8 << 8)]


def test_divide_by_three_trick_is_exact():
    """light4_init divides (sum << 2) by 1..4 (BlockMeshAlgo.hpp:160-161); the kernel uses (n*683)>>11 for /3."""
    n = np.arange(0, 1024, dtype=np.int64)
    assert ((n * 683) >> 11 == n // 3).all()


def test_shard_partition_is_exact():
    from hypercraft_b200 import shard, worldgen

    meshed, resident = worldgen.load_radius_positions(6)
    for world in (1, 2, 4, 8):
        parts = [shard.shard_tasks(meshed, world, r) for r in range(world)]
        allp = np.concatenate(parts)
        assert len(allp) == len(meshed) and len({tuple(p) for p in allp.tolist()}) == len(meshed)
        biggest_plane = int(np.bincount(meshed[:, 0] - meshed[:, 0].min()).max())  # slabs are plane-granular
        assert max(len(p) for p in parts) <= len(meshed) / world + biggest_plane
        for p in parts:
            res = {tuple(q) for q in shard.resident_for(p).tolist()}
            for t in p[:: max(1, len(p) // 20)]:
                assert all(tuple(q) in res for q in worldgen.neighbourhood(t).tolist())
            origin, extent = shard.window_for(shard.resident_for(p))
            assert all(e >= 3 for e in extent)


def test_shard_balance_by_surface_weight():
    """slabs balanced by estimated surface chunks, not by count -- at r = 32 over 8 GPUs the busiest slab holds
    9 % more surface-band chunks than the mean with weights, 35 % more without."""
    from hypercraft_b200 import shard, worldgen

    meshed, _ = worldgen.load_radius_positions(32)
    w = shard.surface_weights(meshed)

    def worst(weights):
        parts = [shard.shard_tasks(meshed, 8, r, weights) for r in range(8)]
        assert sum(len(p) for p in parts) == len(meshed)
        surf = [int((np.abs(p[:, 1]) <= 3).sum()) for p in parts]
        return max(surf) / (sum(surf) / 8)

    assert worst(w) < 1.12 < 1.3 < worst(None)


def test_library_exports_every_declared_symbol():
    """libhcm.so loads and exports exactly the entry points include/hcm.h declares (no compute call)."""
    from hypercraft_b200 import mesher

    if not os.path.exists(mesher.LIB_PATH):
        mesher.build_library()
    lib = ctypes.CDLL(mesher.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "hcm.h")).read()
    declared = sorted(set(re.findall(r"\b(hcm_[a-z_0-9]+)\s*\(", header)))
    assert set(declared) == set(mesher.EXPORTS), (declared, mesher.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"libhcm.so does not export {name}"
    lib.hcm_status_string.restype = ctypes.c_char_p
    assert lib.hcm_status_string(0) == b"ok"


def test_struct_sizes_match_header():
    from hypercraft_b200 import mesher, registry

    assert registry.PROP_DTYPE.itemsize == 20 and registry.FACE_DTYPE.itemsize == 24 and registry.MESH_DTYPE.itemsize == 400
    assert mesher.MESH_DESC_DTYPE.itemsize == 56 and mesher.TASK_RESULT_DTYPE.itemsize == 8
    assert ctypes.sizeof(mesher._Config) == 32 and ctypes.sizeof(mesher._Registry) == 48


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device hcm_create returns HCM_ERR_CUDA."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hypercraft_b200 import mesher

    with pytest.raises(mesher.HcmError) as e:
        mesher.Mesher(origin=(0, 0, 0), extent=(3, 3, 3))
    assert e.value.status == -2 and "no CPU fallback" in str(e.value)


def test_rejects_bad_registry_blob():
    from hypercraft_b200 import mesher

    lib = mesher.load_library()
    reg = mesher._Registry()
    bad = np.zeros(64, dtype=np.uint8)
    assert lib.hcm_registry_from_blob(bad.ctypes.data_as(ctypes.c_void_p), bad.size, ctypes.byref(reg)) == -1


def test_synthetic_chunks_are_deterministic():
    from hypercraft_b200 import worldgen

    a = worldgen.synth_chunk("mix", 3, (1, -2, 0))
    b = worldgen.synth_chunk("mix", 3, (1, -2, 0))
    c = worldgen.synth_chunk("mix", 3, (1, -2, 1))
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and not (a[0] == c[0]).all()
    import hashlib

    assert hashlib.sha256(a[0].tobytes()).hexdigest()[:16] == hashlib.sha256(b[0].tobytes()).hexdigest()[:16]


def test_edit_fan_out_matches_reference_rule():
    """ChunkShouldRemesh / remesh fan-out (ChunkMeshTask.inl:9-15, ChunkSetBlockTask.cpp:100-107)."""
    import edit_oracle as edits

    assert edits.chunk_should_remesh((16, 16, 16)) and edits.chunk_should_remesh((-15, 0, 0)) and not edits.chunk_should_remesh((-16, 0, 0))
    assert edits.chunk_should_remesh((33 + 14, 5, 5)) and not edits.chunk_should_remesh((33 + 15, 5, 5))
    assert not edits.chunk_should_remesh((-8, -8, 40))  # 7 + 7 + 7 = 21
    centre = edits.remesh_offsets((16, 16, 16))
    assert centre == [(0, 0, 0)]  # 16 voxels from every face: no neighbour within 14 of its [-1,33] box
    corner = edits.remesh_offsets((0, 0, 0))
    assert (0, 0, 0) in corner and (-1, -1, -1) in corner and (1, 0, 0) not in corner and len(corner) == 8
    edge = edits.remesh_offsets((31, 10, 16))
    assert (1, 0, 0) in edge and (1, -1, 0) in edge and (-1, 0, 0) not in edge


def test_column_sunlight_height_rule():
    import edit_oracle as edits
    from hypercraft_b200 import registry

    reg = registry.load()
    vpass = reg.props["transparency"][reg.prop_index] == 2
    b = np.zeros(32768, dtype=np.uint16)
    assert edits.column_sunlight_height(b, 3, 4, 0, vpass) == 0
    b[3 + (10 * 32 + 4) * 32] = 1  # stone at y = 10
    assert edits.column_sunlight_height(b, 3, 4, 0, vpass) == 11
    b[3 + (20 * 32 + 4) * 32] = 14  # water (semi-transparent) at y = 20 blocks vertical light too
    assert edits.column_sunlight_height(b, 3, 4, 0, vpass) == 21
    assert edits.column_sunlight_height(b, 3, 4, 5, vpass) == 32


def test_apply_block_edits_chain():
    """The settled state of kSetBlock -> kFloodSunlight -> kSetSunlight -> kFloodSunlight(below): opening and closing a
    shaft through three stacked chunks, no-op edits, last-update-wins, and the remesh fan-out of a height change."""
    import edit_oracle as edits
    from hypercraft_b200 import registry

    reg = registry.load()
    vpass = reg.props["transparency"][reg.prop_index] == 2
    chunks = {}
    for cy in range(-1, 3):  # column of chunks at cx = cz = 0; y = 2 is open sky, the others solid stone
        b = np.full(32768, 0 if cy == 2 else 1, dtype=np.uint16)
        h = np.full(1024, 0 if cy == 2 else 32, dtype=np.int8)
        chunks[(0, cy, 0)] = (b, h)
    x, z = 5, 7
    xz = x + z * 32
    idx = lambda y: x + (y * 32 + z) * 32
    # no-op: stone over stone changes nothing and floods nothing
    n, sun, remesh = edits.apply_block_edits(chunks, [((0, 1, 0), idx(31), 1)], vpass)
    assert n == 0 and not sun and not remesh
    # dig the whole column of chunk y=1 and the top 5 voxels of chunk y=0 in one batch
    batch = [((0, 1, 0), idx(y), 0) for y in range(32)] + [((0, 0, 0), idx(y), 0) for y in range(27, 32)]
    batch.append(((0, 0, 0), idx(27), 1))  # the same voxel again: the later update wins, y = 27 stays stone
    n, sun, remesh = edits.apply_block_edits(chunks, batch, vpass)
    assert n == 32 + 4
    assert sun == {((0, 1, 0), xz): (32, 0), ((0, 0, 0), xz): (32, 28)}
    assert int(chunks[(0, 1, 0)][1][xz]) == 0 and int(chunks[(0, 0, 0)][1][xz]) == 28 and int(chunks[(0, -1, 0)][1][xz]) == 32
    assert (1, 2, 1) in remesh and (-1, -1, -1) in remesh and (0, -2, 0) not in remesh  # 27 around both changed chunks
    # close the shaft at the very top: both chunks below fall back to "not lit from above"
    n, sun, remesh = edits.apply_block_edits(chunks, [((0, 1, 0), idx(31), 1)], vpass)
    assert n == 1 and sun == {((0, 1, 0), xz): (0, 32), ((0, 0, 0), xz): (28, 32)}
    # the top chunk has nothing above it in the map: its chain cannot start (Pop needs the chunk above)
    n, sun, remesh = edits.apply_block_edits(chunks, [((0, 2, 0), idx(3), 1)], vpass)
    assert n == 1 and not sun and (0, 2, 0) in remesh


@pytest.mark.parametrize("force_port", [False, True])
def test_bench_reference_arm_prints_the_contract_line(force_port):
    """`bench.py --impl reference` on a small sphere: oracle/_ref when it is built, the oracle port otherwise
    (HCM_BENCH_FORCE_PORT selects the port to cover that leg where _ref exists).  Needs no GPU."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.pop("HCM_BENCH_FORCE_PORT", None)
    if force_port:
        env["HCM_BENCH_FORCE_PORT"] = "1"
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--radius", "3", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "1"], env=env, capture_output=True, text=True, timeout=300, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    have_ref = os.path.exists(os.path.join(root, "oracle", "_ref", "libhcref.so"))
    assert line["impl"] == "reference" and line["metric"] == "chunks meshed/sec" and line["unit"] == "chunks/s"
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref and not force_port else "port")
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["chunks_per_step"] == 123  # the whole r=3 sphere (misc/load_radius.cpp counts)


def test_port_oracle_batch_equals_single_calls(port):
    """PortOracle.mesh_batch (bench.py's "port" CPU arm) returns what one mesh() call per chunk returns."""
    from hypercraft_b200 import worldgen

    meshed, resident = worldgen.load_radius_positions(1)
    chunks = {tuple(int(c) for c in p): worldgen.synth_terrain_chunk(7, p) for p in resident}
    st, per, bufs = port.mesh_batch(chunks, meshed, threads=3, keep=True)
    assert st["chunks"] == len(meshed) and st["vertices"] == int(per[:, 1].sum()) > 0
    for i, p in enumerate(meshed):
        want = port.mesh(chunks, p)
        assert bufs[i].size == want.size and (bufs[i] == want).all()


def test_digest_batch_is_the_cpu_arms_per_chunk_digest(port, have_ref):
    """bench.py's parity gate compares EVERY chunk through hcoracle.digest_batch: rows built from an hcm_batch_view must equal the
    rows the CPU arms produce for the same chunks (port: PortOracle.mesh_batch; reference: hcref_mesh_batch's per_chunk), and a
    single flipped vertex bit must change exactly that chunk's digest."""
    import hcoracle
    from hypercraft_b200 import worldgen
    from hypercraft_b200.mesher import MESH_DESC_DTYPE, TASK_RESULT_DTYPE, BatchResult

    meshed, resident = worldgen.load_radius_positions(2)
    chunks = {tuple(int(c) for c in p): worldgen.synth_terrain_chunk(9, p) for p in resident}
    st, per, bufs = port.mesh_batch(chunks, meshed, threads=3, keep=True)
    assert per.shape == (len(meshed), 4) and int(per[:, 1].sum()) > 0
    tasks = np.zeros(len(meshed), dtype=TASK_RESULT_DTYPE)
    descs, verts, idxs = [], [np.zeros(3, dtype=np.uint64)], [np.zeros(5, dtype=np.uint16)]  # arenas do not start at a task's data
    nv, ni = 3, 5
    for t, buf in enumerate(bufs):
        assert hcoracle.fnv1a(buf) == int(per[t, 3])
        ms = hcoracle.parse_meshes(buf)
        tasks[t] = (len(descs), len(ms))
        for m in ms:
            d = np.zeros((), dtype=MESH_DESC_DTYPE)
            d["task"], d["transparent"], d["vertex_offset"], d["index_offset"] = t, m["transparent"], nv, ni
            d["vertex_count"], d["index_count"], d["aabb_min"], d["aabb_max"] = len(m["vertices"]), len(m["indices"]), m["aabb_min"], m["aabb_max"]
            descs.append(d)
            verts.append(m["vertices"]), idxs.append(m["indices"])
            nv, ni = nv + len(m["vertices"]), ni + len(m["indices"])
    res = BatchResult(tasks, np.array(descs, dtype=MESH_DESC_DTYPE), np.concatenate(verts), np.concatenate(idxs))
    got = hcoracle.digest_batch(res, 3)
    assert (got == per).all()
    heavy = int(np.argmax(per[:, 1]))
    res.vertices[int(res.meshes[int(tasks[heavy]["first_mesh"])]["vertex_offset"]) + 1] ^= np.uint64(1 << 40)
    bad = (hcoracle.digest_batch(res, 1) != per).any(axis=1)
    assert bad.sum() == 1 and bad[heavy]
    if have_ref:
        ref = hcoracle.RefWorld()
        for p, (b, h) in chunks.items():
            ref.set_chunk(p, b, h)
        _, rper = ref.mesh_batch(meshed, threads=2, per_chunk=True)
        assert (rper == per).all()  # the port is pinned to the reference, digest for digest
        assert hcoracle.fnv1a(ref.mesh(meshed[heavy])) == int(rper[heavy, 3])


def test_cpp_host_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    """hypercraft_b200/host/hcm_host.hpp compiles warning-free as C++20 against include/hcm.h (what a HyperCraft build does) and,
    on a machine without a GPU, MeshTaskDispatcher's constructor throws hcm_host::Error -- it never meshes on the CPU."""
    import subprocess

    import torch

    exe = str(tmp_path / "test_dispatcher")
    lib = os.path.join(ROOT, "hypercraft_b200", "csrc")
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-Wall", "-Wextra", "-Werror", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_dispatcher.cpp"),
                        f"-L{lib}", "-lhcm", "-ldl", f"-Wl,-rpath,{lib}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is tests/test_gpu_parity.py::test_cpp_host_mirror_dispatcher")
    r = subprocess.run([exe, os.path.join(ROOT, "hypercraft_b200", "data", "registry.bin"), os.path.join(ROOT, "oracle", "libhcoracle.so")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr and "PASS" not in r.stdout


def test_bench_algorithmic_byte_figures():
    """canonical 34^3*3 + 8V per chunk, "full" 488 188 + 8V, "min" 66 560 + 11V + 40M."""
    import sys

    sys.path.insert(0, ROOT)
    import bench

    a = bench.algorithmic_bytes(1, 655, 2)
    assert a == {"canonical": 117912 + 8 * 655, "full": 488188 + 8 * 655, "min": 66560 + 11 * 655 + 80}
    # the r = 20 sphere of profiles/r01_head_bench_r20_n1.json
    assert bench.algorithmic_bytes(33401, 32818052, 0)["canonical"] == 4200923128


def test_bench_cpu_side_helpers(have_ref):
    """The pieces of bench.py's b200 arm that need no GPU: secondary roofline figures, the CPU arm at the reference's worker count
    (on oracle/_ref when present and on the port), JSON-serialisable as they stand."""
    import json
    import sys

    sys.path.insert(0, ROOT)
    import bench
    from hypercraft_b200 import worldgen

    sec = bench.roofline_secondary(bench.algorithmic_bytes(33401, 32818052, 11000), 2.488064, 6541.8)
    assert abs(sec["secondary"]["full"]["frac"] - 1.018) < 2e-3 and abs(sec["secondary"]["min"]["achieved"] - 1039) < 2
    meshed, resident = worldgen.load_radius_positions(1)
    arms = [bench.PortArm({tuple(int(c) for c in p): worldgen.synth_terrain_chunk(3, p) for p in resident})]
    if have_ref:
        import hcoracle

        ref = hcoracle.RefWorld()
        ref.generate(resident)
        arms.append(ref)
    for arm in arms:
        st, per = arm.mesh_batch(meshed, threads=2, per_chunk=True)
        assert st["chunks"] == len(meshed) and per.shape[0] == len(meshed) and per.shape[1] >= 3
        assert arm.mesh(meshed[0]).dtype == np.uint8
        pb = bench.pool_baseline(arm, meshed, 8)
        assert pb["cores"] == 6 and pb["value"] > 0
        json.dumps({**sec, "pool": pb, "algo": bench.algorithmic_bytes(3, 4, 5)})


@pytest.mark.parametrize("name", ["test_request_queue", "test_dispatcher_logic", "test_sharded_dispatcher"])
def test_cpp_host_mirror_logic(tmp_path, name):
    """The C++ host mirror's own logic, no GPU and no libhcm.so: the request queue's dedupe / priority rules
    (ChunkMeshTask.inl:38-48) and MeshTaskDispatcher against a recording fake of the C ABI (which chunks cross the
    boundary, one task per position, high priority first, task budget, task -> sink mapping), and ShardedMeshDispatcher
    against one fake ctx per GPU (tasks on the slab owner, boundary planes on both GPUs, edits next to a cut)."""
    import subprocess

    exe = str(tmp_path / name)
    r = subprocess.run(["g++", "-std=c++20", "-O1", "-Wall", "-Wextra", "-Werror", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "cpp", name + ".cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1].startswith("PASS"), r.stdout + r.stderr


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_edit_round_equals_the_undivided_world(port, world):
    """SURVEY 8e / config 5 on N ranks, host logic only (apply_block_edits stands in for hcm_set_blocks, the CPU oracle
    for the mesh kernel): edits go to the slab owner, touched chunks are mirrored to the ranks that hold them as halo,
    remesh tasks go to the rank that owns their position -- block changes, sunlight changes, remesh set and every
    remeshed chunk's bytes equal those of one world taking the same edits."""
    import hashlib

    import edit_oracle as edits
    from hypercraft_b200 import shard, worldgen

    meshed, resident = worldgen.load_radius_positions(3)
    vpass = worldgen._vertical_pass_table()
    bounds = shard.slab_bounds(meshed, world)
    base = {tuple(int(c) for c in p): worldgen.synth_terrain_chunk(5, p) for p in resident}

    def clone(keys):
        return {k: (base[k][0].copy(), base[k][1].copy()) for k in keys}

    # seeded edits in task chunks on both sides of every cut and in the middle: dig, build towers (sunlight), lamps, water
    rng = np.random.default_rng(42)
    task_set = {tuple(int(c) for c in p) for p in meshed}
    cut_planes = sorted({b[0] for b in bounds[1:]} | {b[0] - 1 for b in bounds[1:]} | {0})
    cands = [t for t in sorted(task_set) if t[0] in cut_planes and abs(t[1]) <= 1]
    edit_list = []
    for k in range(60):
        c = cands[int(rng.integers(0, len(cands)))]
        x, y, z = (int(v) for v in rng.integers(0, 32, 3))
        if k % 4 == 0:
            x = int(rng.choice([0, 1, 30, 31]))  # next to the slab face: the fan-out crosses the cut
        block = [worldgen.AIR, worldgen.STONE, worldgen.GLOWSTONE, worldgen.WATER][k % 4]
        edit_list.append((c, x + (y * 32 + z) * 32, block))

    one = clone(base)
    n1, sun1, remesh1 = edits.apply_block_edits(one, edit_list, vpass)
    remesh1 = {p for p in remesh1 if p in task_set}
    assert n1 > 30 and len(sun1) > 5 and len(remesh1) > 10

    res = [shard.resident_for_slab(resident, bounds, r) for r in range(world)]
    for r in range(world):  # a superset of the bench's minimal residency, inside the world
        mine = {tuple(int(c) for c in p) for p in res[r]}
        assert {tuple(int(c) for c in p) for p in shard.resident_for(shard.shard_tasks(meshed, world, r))} <= mine <= set(base)
    worlds = [clone(tuple(int(c) for c in p) for p in res[r]) for r in range(world)]
    blind = [clone(tuple(int(c) for c in p) for p in res[r]) for r in range(world)]  # control: owners edit, nobody mirrors
    n_sum, sun_all, remesh_all, mirrored = 0, {}, set(), 0
    for r, mine in enumerate(shard.route_edits(edit_list, bounds)):
        edits.apply_block_edits(blind[r], mine, vpass)
        n, sun, remesh = edits.apply_block_edits(worlds[r], mine, vpass)
        n_sum += n
        assert not set(sun) & set(sun_all)
        sun_all.update(sun)
        lo, hi = bounds[r]
        # what the owner's ctx can report (positions it sees whole: its own slab) + the host-derived fan-out into other slabs
        remesh_all |= {p for p in remesh if lo <= p[0] < hi} | shard.foreign_fanout(mine, {c for c, _ in sun}, bounds, r)
        touched = {e[0] for e in mine} | {c for c, _ in sun}
        for h, chunks in enumerate(shard.mirror_plan(touched, bounds)):
            for c in chunks:  # the halo holder takes the owner's chunk as data
                assert h != r and c in worlds[h]
                worlds[h][c] = (worlds[r][c][0].copy(), worlds[r][c][1].copy())
                mirrored += 1
    remesh_all = {p for p in remesh_all if p in task_set}
    assert n_sum == n1 and sun_all == sun1 and remesh1 <= remesh_all  # a superset only by edits that changed nothing
    assert any(shard.owner_rank(p, bounds) != shard.owner_rank(e[0], bounds) for e in edit_list for p in [tuple(np.add(e[0], d)) for d in edits.remesh_offsets((e[1] & 31, e[1] >> 10, (e[1] >> 5) & 31))])
    stale = 0
    for r, tasks in enumerate(shard.split_tasks(remesh_all, bounds)):
        lo, hi = bounds[r]
        for t in tasks:
            assert lo <= t[0] < hi
            a, b = port.mesh(worlds[r], t), port.mesh(one, t)
            assert hashlib.sha256(a.tobytes()).digest() == hashlib.sha256(b.tobytes()).digest(), (r, t)
            c = port.mesh(blind[r], t)
            stale += not (c.size == b.size and (c == b).all())
    assert mirrored > 0 and stale > 0  # the cuts were really crossed: without the mirror step some meshes come out stale


def test_bench_watchdog_prints_the_line_and_leaves():
    """bench.py's watchdog: when a stage outlives its budget, rank 0 prints the line built from what is complete (with the stage named under
    "watchdog") and the process exits 0; without a headline yet it exits 3; other ranks print nothing."""
    import json
    import subprocess
    import sys

    prog = ("import sys, time; sys.path.insert(0, %r); import bench; wd = bench.Watchdog(int(sys.argv[1]));"
            "wd.line_fn = (lambda: {'metric': 'chunks meshed/sec', 'value': 1.0}) if sys.argv[2] == 'headline' else None;"
            "wd.arm('optional block', 0.5); time.sleep(30); print('not reached')") % ROOT
    r = subprocess.run([sys.executable, "-c", prog, "0", "headline"], capture_output=True, text=True, timeout=60)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert r.returncode == 0 and line["value"] == 1.0 and "optional block" in line["watchdog"] and "not reached" not in r.stdout
    r = subprocess.run([sys.executable, "-c", prog, "0", "none"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3 and r.stdout.strip() == ""
    r = subprocess.run([sys.executable, "-c", prog, "1", "headline"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_guarded_turns_exceptions_into_an_error_entry():
    import bench

    assert bench.guarded("x", lambda: {"ok": 1}) == {"ok": 1}
    out = bench.guarded("device sink", lambda: 1 / 0)
    assert "error" in out and "ZeroDivisionError" in out["error"] and out["error"].startswith("device sink")


def test_bench_core_split_follows_the_gpus_numa_nodes():
    """Ranks sharing a host: a rank gets cores of its GPU's NUMA node (shared in rank order among the ranks of that node), an even split
    of the mask where the node is unknown; nobody ends up without a core."""
    import bench

    assert bench.parse_cpulist("0-3,8-9\n") == [0, 1, 2, 3, 8, 9] and bench.parse_cpulist("5") == [5]
    allowed = list(range(32))
    node0, node1 = list(range(0, 16)), list(range(16, 32))
    got = bench.split_cores_by_gpu(allowed, [node0, node0, node0, node0, node1, node1, node1, node1])
    assert got[0] == [0, 1, 2, 3] and got[3] == [12, 13, 14, 15] and got[4] == [16, 17, 18, 19] and got[7] == [28, 29, 30, 31]
    assert sorted(c for g in got for c in g) == allowed
    # unknown nodes: even split; a mask smaller than the node list is respected
    assert bench.split_cores_by_gpu(allowed, [None] * 4) == [list(range(0, 8)), list(range(8, 16)), list(range(16, 24)), list(range(24, 32))]
    got = bench.split_cores_by_gpu(list(range(8)), [node0, node1])
    assert got[0] == list(range(8)) and got[1] == list(range(4, 8))  # node 1 has no allowed core: even split for that rank
    # more ranks on a node than it has allowed cores: fall back, never an empty set
    assert all(bench.split_cores_by_gpu([0, 1], [node0] * 4))
