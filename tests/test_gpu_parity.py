"""GPU parity: the CUDA mesher (through the C ABI) against the CPU oracle, byte for byte.

Bar (BASELINE.json north_star): bit-exact per chunk on mesh count, transparent flags, AABBs, vertex bytes and index
bytes.  Inputs are seeded synthetic chunks (machine independent: their meshes and light volumes are also held to the
digests the UNMODIFIED reference produced, tests/golden/synth_golden.json) and, where oracle/_ref is present, terrain
from the reference's own generator meshed by the unmodified reference mesher."""
import hashlib
import json
import os

import numpy as np
import pytest

from util import describe_diff, synth_world

pytestmark = pytest.mark.gpu

# digests of the UNMODIFIED reference's output on the same synthetic inputs (tests/golden/make_golden.py)
GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_golden.json")))

KINDS = ["air", "stone", "checker", "transparent", "mix", "rawmeta", "caves", "flow"]


@pytest.fixture(scope="module")
def mesher():
    from hypercraft_b200 import mesher as M

    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    yield m
    m.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("seed", [1, 2])
def test_light_stage_matches_oracle(mesher, port, kind, seed):
    pos, blocks, heights, chunks = synth_world(kind, seed)
    mesher.upload(pos, blocks, heights)
    got = mesher.debug_light((0, 0, 0))
    want = port.light34(chunks, (0, 0, 0))
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, f"{len(bad)} light bytes differ, first at {bad[0]}: {got[bad[0]]:#x} vs {want[bad[0]]:#x}"
    assert hashlib.sha256(got.tobytes()).hexdigest() == GOLDEN["cases"][f"{kind}/{seed}"]["light_sha256"]  # the reference's own bytes


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_mesh_matches_oracle(mesher, port, kind, seed):
    pos, blocks, heights, chunks = synth_world(kind, seed)
    mesher.upload(pos, blocks, heights)
    res = mesher.mesh(np.array([[0, 0, 0]], dtype=np.int16))
    got = res.serialise(0)
    want = port.mesh(chunks, (0, 0, 0))
    assert got.size == want.size and (got == want).all(), describe_diff(got, want)
    assert hashlib.sha256(got.tobytes()).hexdigest() == GOLDEN["cases"][f"{kind}/{seed}"]["mesh_sha256"]  # the reference's own bytes


TERRAIN = [(12314524, (0, 0, 0)), (12314524, (0, 1, 2)), (12314524, (2, -1, 1)), (7, (-2, 0, -1)), (7, (0, 0, 1)), (7, (-1, -1, -2))]


@pytest.mark.parametrize("i", range(len(TERRAIN)))
def test_terrain_goldens(port, i):
    """Surface chunks of the repo's integer-hash height-field terrain (oceans, plants, lamps): CUDA mesh and light volume
    against the oracle and against the digests the unmodified reference produced for the same inputs."""
    from hypercraft_b200 import mesher as M
    from hypercraft_b200 import worldgen

    seed, centre = TERRAIN[i]
    pos = worldgen.neighbourhood(centre)
    data = [worldgen.synth_terrain_chunk(seed, p) for p in pos]
    blocks, heights = np.stack([d[0] for d in data]), np.stack([d[1] for d in data])
    chunks = {tuple(int(c) for c in p): d for p, d in zip(pos, data)}
    m = M.Mesher(origin=tuple(c - 2 for c in centre), extent=(5, 5, 5))
    try:
        m.upload(pos, blocks, heights)
        got = m.mesh(np.array([centre], dtype=np.int16)).serialise(0)
        light = m.debug_light(centre)
    finally:
        m.close()
    want = port.mesh(chunks, centre)
    assert got.size == want.size and (got == want).all(), describe_diff(got, want)
    assert (light == port.light34(chunks, centre)).all()
    assert hashlib.sha256(got.tobytes()).hexdigest() == GOLDEN["terrain"][i]["mesh_sha256"]
    assert hashlib.sha256(light.tobytes()).hexdigest() == GOLDEN["terrain"][i]["light_sha256"]


def test_mixed_neighbourhood(mesher, port):
    """Centre of one kind, neighbours of others: halo faces, halo flowing water, halo light sources."""
    combos = [("stone", {(1, 0, 0): "air", (0, 1, 0): "mix", (0, 0, -1): "flow"}),
              ("air", {(0, -1, 0): "stone", (1, 0, 0): "mix"}),
              ("flow", {(1, 0, 0): "mix", (-1, 0, 0): "flow", (0, 0, 1): "stone", (0, 0, -1): "rawmeta"}),
              ("caves", {(0, 1, 0): "air", (0, -1, 0): "stone", (1, 1, 0): "mix"})]
    for seed, (kind, by_off) in enumerate(combos):
        pos, blocks, heights, chunks = synth_world(kind, 10 + seed, kinds_by_offset=by_off)
        mesher.upload(pos, blocks, heights)
        res = mesher.mesh(np.array([[0, 0, 0]], dtype=np.int16))
        got, want = res.serialise(0), port.mesh(chunks, (0, 0, 0))
        assert got.size == want.size and (got == want).all(), f"{kind} {by_off}\n" + describe_diff(got, want)


def test_batch_of_many_tasks_and_reupload(port):
    """A 3x3x3 block of tasks in one batch inside a 5^3 resident window; then edit one chunk and remesh."""
    from hypercraft_b200 import mesher as M, worldgen

    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    allpos = np.array([[x, y, z] for x in range(-2, 3) for y in range(-2, 3) for z in range(-2, 3)], dtype=np.int16)
    kinds = ["caves", "mix", "air", "stone", "transparent", "flow"]
    blocks = np.zeros((len(allpos), 32768), dtype=np.uint16)
    heights = np.zeros((len(allpos), 1024), dtype=np.int8)
    chunks = {}
    for i, p in enumerate(allpos):
        blocks[i], heights[i] = worldgen.synth_chunk(kinds[(int(p[0]) * 7 + int(p[1]) * 3 + int(p[2])) % len(kinds)], 5, p)
        chunks[tuple(int(c) for c in p)] = (blocks[i], heights[i])
    m.upload(allpos, blocks, heights)
    tasks = np.array([[x, y, z] for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)], dtype=np.int16)
    res = m.mesh(tasks)
    for t, p in enumerate(tasks):
        got, want = res.serialise(t), port.mesh(chunks, p)
        assert got.size == want.size and (got == want).all(), f"task {p}\n" + describe_diff(got, want)
    # edit: replace the centre chunk, remesh it and a neighbour
    nb, nh = worldgen.synth_chunk("checker", 9, (0, 0, 0))
    chunks[(0, 0, 0)] = (nb, nh)
    m.upload(np.array([[0, 0, 0]], dtype=np.int16), nb[None], nh[None])
    tasks2 = np.array([[0, 0, 0], [1, 0, 0], [0, -1, 1]], dtype=np.int16)
    res = m.mesh(tasks2)
    for t, p in enumerate(tasks2):
        got, want = res.serialise(t), port.mesh(chunks, p)
        assert got.size == want.size and (got == want).all(), f"after edit, task {p}\n" + describe_diff(got, want)
    m.close()


def test_errors(mesher):
    from hypercraft_b200 import mesher as M

    with pytest.raises(M.HcmError) as e:
        mesher.mesh(np.array([[2, 0, 0]], dtype=np.int16))  # window border: neighbour cannot be resident
    assert e.value.status == -4
    with pytest.raises(M.HcmError) as e:
        mesher.mesh(np.array([[40, 0, 0]], dtype=np.int16))
    assert e.value.status == -5
    m2 = M.Mesher(origin=(0, 0, 0), extent=(3, 3, 3))
    with pytest.raises(M.HcmError) as e:
        m2.mesh(np.array([[1, 1, 1]], dtype=np.int16))  # nothing uploaded
    assert e.value.status == -4
    m2.close()


def test_reference_terrain(have_ref, port):
    """Reference terrain (seed 12314524) radius 3: every meshed chunk equals the unmodified reference mesher."""
    if not have_ref:
        pytest.skip("oracle/_ref/libhcref.so not present")
    import hcoracle
    from hypercraft_b200 import mesher as M, worldgen

    meshed, resident = worldgen.load_radius_positions(3)
    ref = hcoracle.RefWorld()
    ref.generate(resident)
    blocks, heights = ref.get_chunks(resident)
    m = M.Mesher(origin=(-4, -4, -4), extent=(9, 9, 9))
    m.upload(resident, blocks, heights)
    res = m.mesh(meshed)
    n_nonempty = 0
    for t, p in enumerate(meshed):
        got, want = res.serialise(t), ref.mesh(p)
        assert got.size == want.size and (got == want).all(), f"terrain chunk {p}\n" + describe_diff(got, want)
        n_nonempty += int(res.tasks[t]["n_meshes"]) > 0
    assert n_nonempty > 10
    m.close()


def test_cpp_host_mirror_dispatcher():
    """hypercraft_b200/host/hcm_host.hpp (the C++ side a HyperCraft build would compile in): Submit 27-chunk arrays,
    Flush into a PushChunkMesh-shaped sink, compare with the oracle -- see tests/cpp/test_dispatcher.cpp."""
    import subprocess

    import hcoracle

    exe, root = _build_cpp("test_dispatcher")
    hcoracle.build(ref=False, port=True)
    r = subprocess.run([exe, os.path.join(root, "hypercraft_b200", "data", "registry.bin"), hcoracle.PORT_SO], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1].startswith("PASS"), r.stdout + r.stderr


def _build_cpp(name):
    """tests/cpp/<name>.cpp linked against the library under test (libhcm.so; the CPU emulation build under HCM_TEST_SIM)."""
    import subprocess

    from hypercraft_b200 import mesher as M

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", name)
    libdir, libfile = os.path.split(M.LIB_PATH)
    subprocess.run(["g++", "-std=c++20", "-O2", "-pthread", "-o", exe, exe + ".cpp", f"-L{libdir}", "-l" + libfile[3:-3], "-ldl", f"-Wl,-rpath,{libdir}"],
                   check=True)
    return exe, root


def test_cpp_sharded_dispatcher_on_two_devices():
    """hcm_host::ShardedMeshDispatcher::FlushParallel over two hcm_ctx on two devices: world load and edits across the
    cut, every pushed mesh against the oracle -- tests/cpp/test_sharded_gpu.cpp.  Skipped on a one-GPU box."""
    import subprocess

    import hcoracle

    exe, root = _build_cpp("test_sharded_gpu")
    hcoracle.build(ref=False, port=True)
    env = dict(os.environ)
    if os.environ.get("HCM_TEST_SIM"):
        env["CUSIM_DEVICES"] = "2"
    r = subprocess.run([exe, os.path.join(root, "hypercraft_b200", "data", "registry.bin"), hcoracle.PORT_SO], capture_output=True, text=True, env=env)
    last = (r.stdout.strip().splitlines() or [""])[-1]
    if r.returncode == 0 and last.startswith("SKIP"):
        pytest.skip(last)
    assert r.returncode == 0 and last.startswith("PASS"), r.stdout + r.stderr


def _terrain_world(radius):
    """(meshed, resident, blocks, heights): reference terrain if oracle/_ref is present, else the repo's fallback terrain."""
    import hcoracle
    from hypercraft_b200 import worldgen

    meshed, resident = worldgen.load_radius_positions(radius)
    if hcoracle.have_ref():
        ref = hcoracle.RefWorld()
        ref.generate(resident)
        blocks, heights = ref.get_chunks(resident)
    else:
        blocks = np.zeros((len(resident), 32768), dtype=np.uint16)
        heights = np.zeros((len(resident), 1024), dtype=np.int8)
        for i, p in enumerate(resident):
            blocks[i], heights[i] = worldgen.synth_terrain_chunk(12314524, p)
    return meshed, resident, blocks, heights


def test_structural_invariants_at_scale():
    """Size-independent properties on a whole load sphere (r = 8, 2 109 tasks): vertex/index/quad bookkeeping, the
    16-bit cut, index ranges, AABBs that bound their vertices, idempotence and independence from task order."""
    from hypercraft_b200 import mesher as M

    meshed, resident, blocks, heights = _terrain_world(8)
    m = M.Mesher(origin=(-9, -9, -9), extent=(19, 19, 19), max_chunks=len(resident), max_tasks=len(meshed))
    m.upload(resident, blocks, heights)
    res = m.mesh(meshed)
    md = res.meshes
    assert int(md["vertex_count"].sum()) == len(res.vertices) and int(md["index_count"].sum()) == len(res.indices)
    assert (md["vertex_count"] % 4 == 0).all() and (md["index_count"] * 4 == md["vertex_count"] * 6).all()
    assert (md["vertex_count"] <= 65532).all() and (md["vertex_count"] > 0).all()
    assert int(res.tasks["n_meshes"].sum()) == len(md)
    snap = {}
    for t in range(len(meshed)):
        ms = res.task_meshes(t)
        for k, mesh in enumerate(ms):
            v = mesh["vertices"]
            w0 = (v & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            xyz = np.stack([w0 & 1023, (w0 >> 10) & 1023, (w0 >> 20) & 1023], axis=1)
            assert (xyz.min(axis=0) == mesh["aabb_min"]).all() and (xyz.max(axis=0) == mesh["aabb_max"]).all()
            assert (xyz <= 512).all()
            idx = mesh["indices"].astype(np.int64)
            assert idx.max() < len(v) or (idx[:6].max() == 65535 and idx[6:].max(initial=0) < len(v))  # stale base of a continuation mesh
            assert mesh["transparent"] in (0, 1)
        snap[t] = res.serialise(t).tobytes()
    # idempotence
    res2 = m.mesh(meshed)
    assert all(res2.serialise(t).tobytes() == snap[t] for t in range(0, len(meshed), 7))
    # task order does not matter
    perm = np.random.default_rng(3).permutation(len(meshed))
    res3 = m.mesh(meshed[perm])
    assert all(res3.serialise(i).tobytes() == snap[int(perm[i])] for i in range(0, len(meshed), 5))
    # re-deriving from the device-resident Block data changes nothing
    m.rederive()
    res4 = m.mesh(meshed)
    assert all(res4.serialise(t).tobytes() == snap[t] for t in range(0, len(meshed), 11))
    m.close()


def test_translation_invariance_and_removal(port):
    """The same 27 chunks give the same meshes wherever the window puts them; removed chunks stop being resident."""
    from hypercraft_b200 import mesher as M

    pos, blocks, heights, chunks = synth_world("mix", 4)
    want = port.mesh(chunks, (0, 0, 0))
    for origin, shift in [((-3, -2, -4), (0, 0, 0)), ((100, -50, 7), (102, -47, 9))]:
        m = M.Mesher(origin=origin, extent=(7, 6, 8))
        p2 = (pos.astype(np.int32) + np.array(shift)).astype(np.int16)
        m.upload(p2, blocks, heights)
        res = m.mesh(np.array([shift], dtype=np.int16))
        got = res.serialise(0)
        assert got.size == want.size and (got == want).all()
        assert m.is_resident(shift)
        m.remove(np.array([shift], dtype=np.int16))
        assert not m.is_resident(shift)
        with pytest.raises(M.HcmError) as e:
            m.mesh(np.array([shift], dtype=np.int16))
        assert e.value.status == -4
        s = m.stats()
        assert s["kernel_launches"] >= 3 and s["h2d_bytes"] > 0 and s["d2h_bytes"] > 0 and s["device_bytes"] > 0
        m.close()


def test_output_arena_grows_instead_of_truncating(port):
    """A deliberately tiny vertex arena: the batch is re-run with a larger one, never cut short."""
    from hypercraft_b200 import mesher as M

    pos, blocks, heights, chunks = synth_world("checker", 2)
    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5), vertex_capacity=1024)
    m.upload(pos, blocks, heights)
    res = m.mesh(np.array([[0, 0, 0]], dtype=np.int16))
    got, want = res.serialise(0), port.mesh(chunks, (0, 0, 0))
    assert got.size == want.size and (got == want).all()
    m.close()


def test_edit_driven_incremental_remesh(port):
    """Config 5: seeded single-block edits; after each, re-upload the edited chunk (blocks + recomputed sunlight column),
    remesh exactly the reference's fan-out set and compare every remeshed chunk with the oracle on the updated world."""
    import edit_oracle as edits
    from hypercraft_b200 import mesher as M, registry, worldgen

    reg = registry.load()
    vpass = reg.props["transparency"][reg.prop_index] == 2
    allpos = np.array([[x, y, z] for x in range(-2, 3) for y in range(-2, 3) for z in range(-2, 3)], dtype=np.int16)
    chunks = {}
    for p in allpos:
        b, h = worldgen.synth_terrain_chunk(77, p)
        chunks[tuple(int(c) for c in p)] = (b.copy(), h.copy())
    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    m.upload(allpos, np.stack([chunks[tuple(int(c) for c in p)][0] for p in allpos]), np.stack([chunks[tuple(int(c) for c in p)][1] for p in allpos]))
    rng = np.random.default_rng(11)
    new_blocks = [0, 1, 24, 14, worldgen.make_block(worldgen.FLOWING_WATER, 5), 9, worldgen.make_block(worldgen.LEAVES, 2)]
    for it in range(10):
        bx, bz = int(rng.integers(0, 32)), int(rng.integers(0, 32))
        cpos = (0, int(rng.integers(-1, 1)), 0)
        b, h = chunks[cpos]
        by = int(np.clip(int(h[bx + bz * 32]) + int(rng.integers(-3, 2)), 0, 31)) if it % 3 else int(rng.integers(0, 32))
        b[bx + (by * 32 + bz) * 32] = new_blocks[it % len(new_blocks)]
        up_h = int(chunks[(cpos[0], cpos[1] + 1, cpos[2])][1][bx + bz * 32])
        h[bx + bz * 32] = edits.column_sunlight_height(b, bx, bz, up_h, vpass)
        m.upload(np.array([cpos], dtype=np.int16), b[None], h[None])
        tasks = np.array([[cpos[0] + d[0], cpos[1] + d[1], cpos[2] + d[2]] for d in edits.remesh_offsets((bx, by, bz))], dtype=np.int16)
        tasks = tasks[(np.abs(tasks) <= 1).all(axis=1)]  # tasks need their 26 neighbours inside the 5^3 window
        res = m.mesh(tasks)
        for t, p in enumerate(tasks):
            got, want = res.serialise(t), port.mesh(chunks, p)
            assert got.size == want.size and (got == want).all(), f"edit {it} at {(bx, by, bz)} in {cpos}: task {p}\n" + describe_diff(got, want)
    m.close()


def _small_world():
    from hypercraft_b200 import worldgen

    allpos = np.array([[x, y, z] for x in range(-2, 3) for y in range(-2, 3) for z in range(-2, 3)], dtype=np.int16)
    kinds = ["caves", "mix", "air", "stone", "transparent", "flow", "checker"]
    blocks = np.zeros((len(allpos), 32768), dtype=np.uint16)
    heights = np.zeros((len(allpos), 1024), dtype=np.int8)
    chunks = {}
    for i, p in enumerate(allpos):
        blocks[i], heights[i] = worldgen.synth_chunk(kinds[(int(p[0]) * 5 + int(p[1]) * 3 + int(p[2])) % len(kinds)], 11, p)
        chunks[tuple(int(c) for c in p)] = (blocks[i], heights[i])
    tasks = np.array([[x, y, z] for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)], dtype=np.int16)
    return allpos, blocks, heights, chunks, tasks


def test_pipelined_upload_and_mesh_small(port, monkeypatch):
    """hcm_upload_and_mesh on a 5^3 world cut into 4 upload blocks: oracle parity per task, rows in the caller's task
    order, chunks that were resident before the call, an empty upload, the arena-overflow fallback and the error path."""
    from hypercraft_b200 import mesher as M

    monkeypatch.setenv("HCM_UPLOAD_BLOCK", "32")
    monkeypatch.setenv("HCM_MIN_SUB_BATCH", "1")
    allpos, blocks, heights, chunks, tasks = _small_world()
    order = np.random.default_rng(5).permutation(len(tasks))

    def check(res, tlist, what):
        assert len(res.tasks) == len(tlist)
        for t, p in enumerate(tlist):
            got, want = res.serialise(t), port.mesh(chunks, p)
            assert got.size == want.size and (got == want).all(), f"{what}: task {p}\n" + describe_diff(got, want)

    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    check(m.upload_and_mesh(allpos, blocks, heights, tasks[order]), tasks[order], "fresh world")
    check(m.upload_and_mesh(allpos[:0], blocks[:0], heights[:0], tasks), tasks, "empty upload")
    check(m.upload_and_mesh(allpos[40:90], blocks[40:90], heights[40:90], tasks), tasks, "partial re-upload")
    m.close()
    # half of the world resident beforehand, the other half arrives with the call (in reverse order)
    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    m.upload(allpos[:60], blocks[:60], heights[:60])
    rest = np.arange(60, len(allpos))[::-1]
    check(m.upload_and_mesh(allpos[rest], blocks[rest], heights[rest], tasks), tasks, "half resident")
    m.close()
    # a vertex arena that is far too small: the call falls back to the growing path, nothing is truncated
    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5), vertex_capacity=1024)
    check(m.upload_and_mesh(allpos, blocks, heights, tasks), tasks, "tiny arena")
    m.close()
    # a task whose neighbourhood stays incomplete: error, and the upload is rolled back
    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    with pytest.raises(M.HcmError) as e:
        m.upload_and_mesh(allpos[:100], blocks[:100], heights[:100], tasks)
    assert e.value.status == -4
    assert not any(m.is_resident(p) for p in allpos)
    with pytest.raises(M.HcmError) as e:
        m.upload_and_mesh(allpos, blocks, heights, np.array([[2, 0, 0]], dtype=np.int16))
    assert e.value.status == -4
    check(m.upload_and_mesh(allpos, blocks, heights, tasks), tasks, "after errors")
    m.close()


def test_pipelined_world_load_equals_plain_calls(monkeypatch):
    """A whole load sphere (r = 8) through hcm_upload_and_mesh in ~16 upload blocks and in a shuffled chunk order gives,
    task by task, the bytes of upload + mesh + fetch."""
    from hypercraft_b200 import mesher as M

    meshed, resident, blocks, heights = _terrain_world(8)
    kw = dict(origin=(-9, -9, -9), extent=(19, 19, 19), max_chunks=len(resident), max_tasks=len(meshed))
    m = M.Mesher(**kw)
    m.upload(resident, blocks, heights)
    plain = m.mesh(meshed)
    want = [plain.serialise(t).tobytes() for t in range(len(meshed))]
    m.close()
    monkeypatch.setenv("HCM_UPLOAD_BLOCK", "256")
    monkeypatch.setenv("HCM_MIN_SUB_BATCH", "64")
    m = M.Mesher(**kw)
    res = m.upload_and_mesh(resident, blocks, heights, meshed)
    assert len(res.vertices) == len(plain.vertices) and len(res.indices) == len(plain.indices) and len(res.meshes) == len(plain.meshes)
    assert all(res.serialise(t).tobytes() == want[t] for t in range(len(meshed)))
    launches = m.stats()["kernel_launches"]
    assert launches >= 3 * 8, f"expected a pipeline of many sub-batches, saw {launches} kernel launches"
    perm = np.random.default_rng(9).permutation(len(resident))
    res = m.upload_and_mesh(resident[perm], blocks[perm], heights[perm], meshed)  # everything resident already, re-uploaded shuffled
    assert all(res.serialise(t).tobytes() == want[t] for t in range(len(meshed)))
    m.close()


def test_set_blocks_on_device_matches_host_chain(port):
    """hcm_set_blocks (device-resident edits + sunlight heights on the GPU) against the host restatement of the
    reference's kSetBlock / kFloodSunlight / kSetSunlight chain: changed-block count, every sunlight change, the mesh tasks
    pushed, and -- the state that matters -- the meshes of ALL interior chunks afterwards against the oracle on the
    host-edited world."""
    import edit_oracle as edits
    from hypercraft_b200 import mesher as M, registry, worldgen

    reg = registry.load()
    vpass = reg.props["transparency"][reg.prop_index] == 2
    allpos = np.array([[x, y, z] for x in range(-2, 3) for y in range(-2, 3) for z in range(-2, 3)], dtype=np.int16)
    chunks = {}
    for p in allpos:
        b, h = worldgen.synth_terrain_chunk(78, p)
        chunks[tuple(int(c) for c in p)] = (b.copy(), h.copy())
    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    m.upload(allpos, np.stack([chunks[tuple(int(c) for c in p)][0] for p in allpos]), np.stack([chunks[tuple(int(c) for c in p)][1] for p in allpos]))
    interior = np.array([[x, y, z] for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)], dtype=np.int16)
    rng = np.random.default_rng(21)
    palette = [0, 1, 24, 14, worldgen.make_block(worldgen.FLOWING_WATER, 5), 9, worldgen.make_block(worldgen.LEAVES, 2), 20, 5]

    def run(batch, what):
        want_n, want_sun, want_remesh = edits.apply_block_edits(chunks, batch, vpass)
        r = m.set_blocks(batch)
        assert r["n_changed_blocks"] == want_n, what
        got_sun = {(tuple(int(v) for v in s["chunk"]), int(s["index"])): (int(s["old_height"]), int(s["new_height"])) for s in r["sunlight"]}
        assert got_sun == want_sun, what
        meshable = {p for p in want_remesh if all(abs(v) <= 1 for v in p)}  # the 27 chunks around must be inside the 5^3 window
        assert {tuple(int(v) for v in p) for p in r["remesh"]} == meshable, what
        res = m.mesh(interior)  # every chunk that could have changed, and the ones that must not have
        for t, p in enumerate(interior):
            got, want = res.serialise(t), port.mesh(chunks, p)
            assert got.size == want.size and (got == want).all(), f"{what}: chunk {p}\n" + describe_diff(got, want)

    # 1. a shaft dug from the sky down through two chunks in ONE batch (the chain has to run top-down), then closed again
    x, z = 9, 20
    col = [((0, cy, 0), x + (y * 32 + z) * 32, 0) for cy in (1, 0) for y in range(32)]
    run(col, "shaft")
    run([((0, 1, 0), x + (31 * 32 + z) * 32, 1)], "lid")
    # 2. random batches around the terrain surface: duplicates, no-ops, several edits per column, neighbours across chunk borders
    for it in range(6):
        batch = []
        for _ in range(int(rng.integers(1, 40))):
            cpos = (int(rng.integers(-1, 2)), int(rng.integers(-1, 2)), int(rng.integers(-1, 2)))
            bx, bz = int(rng.integers(0, 32)), int(rng.integers(0, 32))
            h = int(chunks[cpos][1][bx + bz * 32])
            by = int(np.clip(h + int(rng.integers(-4, 3)), 0, 31)) if rng.integers(0, 3) else int(rng.integers(0, 32))
            batch.append((cpos, bx + (by * 32 + bz) * 32, palette[int(rng.integers(0, len(palette)))]))
            if rng.integers(0, 5) == 0:
                batch.append(batch[-1][:2] + (palette[int(rng.integers(0, len(palette)))],))  # same voxel twice: the last one wins
            if rng.integers(0, 4) == 0:  # another edit in the same world column, other chunk
                batch.append(((cpos[0], int(rng.integers(-1, 2)), cpos[2]), bx + (int(rng.integers(0, 32)) * 32 + bz) * 32, 0))
        run(batch, f"random batch {it}")
    # 3. errors: chunk outside the window, chunk not resident
    with pytest.raises(M.HcmError) as e:
        m.set_blocks([((9, 0, 0), 0, 1)])
    assert e.value.status == -5
    m.remove(np.array([[2, 2, 2]], dtype=np.int16))
    with pytest.raises(M.HcmError) as e:
        m.set_blocks([((2, 2, 2), 0, 1)])
    assert e.value.status == -4
    m.close()


def _device_buffer(nbytes):
    """(raw device pointer, read-back function, keep-alive) of a caller-owned device buffer: torch on the GPU box, plain host memory
    under the CPU emulation (where "device memory" is host memory)."""
    if os.environ.get("HCM_TEST_SIM"):
        a = np.full(nbytes, 0xEE, dtype=np.uint8)
        return a.ctypes.data, (lambda: a.copy()), a
    import torch

    t = torch.full((nbytes,), 0xEE, dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize()
    return t.data_ptr(), (lambda: t.cpu().numpy()), t


def test_whole_load_sphere_is_byte_identical_to_the_cpu_mesher(port):
    """EVERY chunk of a load sphere (r = 8: 2 109 tasks, 3 071 resident chunks of the reference's terrain where oracle/_ref is
    present) against the CPU mesher -- the unmodified reference there, else the port -- through the per-chunk digest of the serialised
    meshes that bench.py's parity gate uses (mesh count, flags, AABBs, vertex bytes, index bytes).  Plain path and pipelined path."""
    import hcoracle
    from hypercraft_b200 import mesher as M, worldgen

    meshed, resident = worldgen.load_radius_positions(8)
    if hcoracle.have_ref():
        ref = hcoracle.RefWorld()
        ref.generate(resident)
        blocks, heights = ref.get_chunks(resident)
        _, want = ref.mesh_batch(meshed, per_chunk=True)
    else:
        blocks = np.zeros((len(resident), 32768), dtype=np.uint16)
        heights = np.zeros((len(resident), 1024), dtype=np.int8)
        chunks = {}
        for i, p in enumerate(resident):
            blocks[i], heights[i] = worldgen.synth_terrain_chunk(12314524, p)
            chunks[tuple(int(c) for c in p)] = (blocks[i], heights[i])
        _, want, _ = port.mesh_batch(chunks, meshed, threads=os.cpu_count() or 1)
    m = M.Mesher(origin=(-9, -9, -9), extent=(19, 19, 19), max_chunks=len(resident), max_tasks=len(meshed))
    m.upload(resident, blocks, heights)
    got = hcoracle.digest_batch(m.mesh(meshed))
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} of {len(meshed)} chunks differ, first: task {meshed[bad[0]]} got {got[bad[0]]} want {want[bad[0]]}"
    assert int(want[:, 1].sum()) > 100000  # a real world, not a sphere of air
    res = m.upload_and_mesh(resident, blocks, heights, meshed)
    got = hcoracle.digest_batch(res)
    assert (got == want).all(), "pipelined path differs"
    # hcm_mesh_desc::task is the caller's task index on the pipelined path as well
    for t in range(0, len(meshed), 13):
        first, n = int(res.tasks[t]["first_mesh"]), int(res.tasks[t]["n_meshes"])
        assert (res.meshes["task"][first:first + n] == t).all()
    m.close()


def test_hinted_upload_and_scan_balance(port, monkeypatch):
    """hcm_upload_chunks_hinted: chunks the caller declares uniform are derived from the hint without their 64 KB being read (the
    arrays passed for them hold garbage here), HCM_HINT_NOT_UNIFORM chunks go up unscanned, HCM_HINT_UNKNOWN ones are scanned; and a
    large unhinted upload gives the same meshes whether the scan threads or the copy engine take the work items (HCM_UPLOAD_ADAPTIVE)."""
    from hypercraft_b200 import mesher as M

    allpos, blocks, heights, chunks, tasks = _small_world()
    uniform = (blocks == blocks[:, :1]).all(axis=1)
    assert uniform.sum() > 20 and (~uniform).sum() > 20
    hint = np.where(uniform, blocks[:, 0].astype(np.int32), M.HINT_NOT_UNIFORM).astype(np.int32)
    hint[::3] = np.where(uniform[::3], hint[::3], M.HINT_UNKNOWN)
    poisoned = blocks.copy()
    poisoned[uniform] = 0xBEEF  # must never be looked at
    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    m.upload(allpos, poisoned, heights, uniform_hint=hint)
    res = m.mesh(tasks)
    for t, p in enumerate(tasks):
        got, want = res.serialise(t), port.mesh(chunks, p)
        assert got.size == want.size and (got == want).all(), f"hinted upload: task {p}\n" + describe_diff(got, want)
    with pytest.raises(M.HcmError) as e:
        m.upload(allpos[:1], blocks[:1], heights[:1], uniform_hint=np.array([70000], dtype=np.int32))
    assert e.value.status == -1
    m.close()
    # >= 256 chunks: the threaded path.  Same bytes with the balance on and off; `upload_unscanned_chunks` says who took what.
    meshed, resident, wb, wh = _terrain_world(5)
    kw = dict(origin=(-6, -6, -6), extent=(13, 13, 13), max_chunks=len(resident), max_tasks=len(meshed))
    out = {}
    for adaptive in ("0", "1"):
        monkeypatch.setenv("HCM_UPLOAD_ADAPTIVE", adaptive)
        monkeypatch.setenv("HCM_UPLOAD_BLOCK", "128")
        monkeypatch.setenv("HCM_UPLOAD_FIRST_BLOCK", "64")
        m = M.Mesher(**kw)
        res = m.upload_and_mesh(resident, wb, wh, meshed)
        out[adaptive] = [res.serialise(t).tobytes() for t in range(len(meshed))]
        unscanned = m.stats()["upload_unscanned_chunks"]
        assert unscanned == 0 if adaptive == "0" else unscanned <= len(resident)
        m.close()
    assert out["0"] == out["1"]


def test_device_sink_emit_into(port):
    """SURVEY 8f row 2, the CUDA half: hcm_mesh_counts -> the host allocates a range per mesh the way MeshCluster::try_alloc does
    (client/include/client/mesh/MeshCluster.hpp:38-67; here: shuffled order, odd alignment padding) -> hcm_mesh_emit_into writes the
    vertices and indices into caller-owned DEVICE buffers.  The bytes at every mesh's offsets equal the fetched (host) output, which
    equals the oracle; bytes between the ranges are untouched; the error paths leave the buffers alone."""
    from hypercraft_b200 import mesher as M

    allpos, blocks, heights, chunks, tasks = _small_world()
    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    m.upload(allpos, blocks, heights)
    m.mesh_async(tasks)
    cnt = m.counts()
    assert len(cnt.vertices) == 0 and len(cnt.indices) == 0 and len(cnt.tasks) == len(tasks)
    md = cnt.meshes
    assert int(md["vertex_count"].sum()) == m.n_vertices and int(md["index_count"].sum()) == m.n_indices and len(md) > 20
    # allocator: meshes placed in a shuffled order, each range preceded by a gap of 1..5 vertices / 1..7 indices
    rng = np.random.default_rng(17)
    order = rng.permutation(len(md))
    pl = np.zeros(len(md), dtype=M.MESH_PLACEMENT_DTYPE)
    vpos, ipos = 3, 1
    for k, mi in enumerate(order):
        vpos += int(rng.integers(1, 6))
        ipos += int(rng.integers(1, 8))
        pl[k] = (mi, 0, vpos, ipos)
        vpos += int(md["vertex_count"][mi])
        ipos += int(md["index_count"][mi])
    vcap, icap = vpos + 9, ipos + 5
    vptr, vread, _vkeep = _device_buffer(vcap * 8)
    iptr, iread, _ikeep = _device_buffer(icap * 2)
    m.emit_into(pl, vptr, vcap, iptr, icap)
    assert m.stats()["sink_ms"] >= 0.0
    vbuf, ibuf = vread().view(np.uint64), iread().view(np.uint16)
    host = m.fetch()  # the same batch through the host path
    assert (host.meshes == md).all()
    vmask, imask = np.zeros(vcap, dtype=bool), np.zeros(icap, dtype=bool)
    for k, mi in enumerate(order):
        d = md[mi]
        vo, io, nv, ni = int(pl[k]["vertex_offset"]), int(pl[k]["index_offset"]), int(d["vertex_count"]), int(d["index_count"])
        assert (vbuf[vo:vo + nv] == host.vertices[int(d["vertex_offset"]):int(d["vertex_offset"]) + nv]).all(), f"mesh {mi}: vertices"
        assert (ibuf[io:io + ni] == host.indices[int(d["index_offset"]):int(d["index_offset"]) + ni]).all(), f"mesh {mi}: indices"
        vmask[vo:vo + nv] = True
        imask[io:io + ni] = True
    assert (vbuf[~vmask] == np.uint64(0xEEEEEEEEEEEEEEEE)).all() and (ibuf[~imask] == 0xEEEE).all()  # nothing outside the ranges
    for t, p in enumerate(tasks):  # and the host path is the oracle's
        got, want = host.serialise(t), port.mesh(chunks, p)
        assert got.size == want.size and (got == want).all()
    # errors: a range beyond the capacity, overlapping ranges, an unknown mesh, no descriptors yet
    bad = pl.copy()
    bad[0]["vertex_offset"] = vcap - 1
    with pytest.raises(M.HcmError) as e:
        m.emit_into(bad, vptr, vcap, iptr, icap)
    assert e.value.status == -1
    bad = pl.copy()
    bad[1]["index_offset"] = bad[0]["index_offset"] + 2
    with pytest.raises(M.HcmError) as e:
        m.emit_into(bad, vptr, vcap, iptr, icap)
    assert e.value.status == -1
    bad = pl.copy()
    bad[2]["mesh"] = len(md)
    with pytest.raises(M.HcmError) as e:
        m.emit_into(bad, vptr, vcap, iptr, icap)
    assert e.value.status == -1
    assert (vread().view(np.uint64) == vbuf).all() and (iread().view(np.uint16) == ibuf).all()  # failed calls wrote nothing
    m.mesh_async(tasks[:3])
    with pytest.raises(M.HcmError) as e:
        m.emit_into(pl[:1], vptr, vcap, iptr, icap)
    assert e.value.status == -7
    # after hcm_upload_and_mesh the arenas hold the whole call's meshes: emit_into works on them as well
    res = m.upload_and_mesh(allpos[:0], blocks[:0], heights[:0], tasks)
    pl2 = np.zeros(len(res.meshes), dtype=M.MESH_PLACEMENT_DTYPE)
    pl2["mesh"] = np.arange(len(res.meshes))
    pl2["vertex_offset"] = np.concatenate([[0], np.cumsum(res.meshes["vertex_count"])[:-1]])
    pl2["index_offset"] = np.concatenate([[0], np.cumsum(res.meshes["index_count"])[:-1]])
    m.emit_into(pl2, vptr, vcap, iptr, icap)
    vb2 = vread().view(np.uint64)
    for k in range(0, len(pl2), 5):
        d = res.meshes[k]
        vo, nv = int(pl2[k]["vertex_offset"]), int(d["vertex_count"])
        assert (vb2[vo:vo + nv] == res.vertices[int(d["vertex_offset"]):int(d["vertex_offset"]) + nv]).all()
    m.close()


def test_set_blocks_reports_who_pushes_each_remesh():
    """hcm_edit_result::remesh_flags: the fan-out of a changed block is the reference's high-priority push
    (ChunkSetBlockTask.cpp:119), positions only a changed sunlight height reaches are its low-priority push
    (ChunkSetSunlightTask.cpp:84)."""
    from hypercraft_b200 import mesher as M, worldgen

    allpos = np.array([[x, y, z] for x in range(-2, 3) for y in range(-2, 3) for z in range(-2, 3)], dtype=np.int16)
    b = np.zeros((len(allpos), 32768), dtype=np.uint16)
    h = np.zeros((len(allpos), 1024), dtype=np.int8)  # an empty, fully sunlit world
    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    m.upload(allpos, b, h)
    # stone in the middle of chunk (0,0,0): only that chunk is within reach of the block, but the height change pushes all 27
    r = m.set_blocks([((0, 0, 0), 16 + (16 * 32 + 16) * 32, 1)])
    flags = {tuple(int(v) for v in p): int(f) for p, f in zip(r["remesh"], r["remesh_flags"])}
    assert r["n_changed_blocks"] == 1 and len(r["sunlight"]) >= 1 and len(flags) == 27
    assert flags[(0, 0, 0)] == 3 and all(f == 2 for p, f in flags.items() if p != (0, 0, 0))
    # air back in, next to the -x face: (0,0,0) and (-1,0,0) are pushed by the block edit
    r = m.set_blocks([((0, 0, 0), 16 + (16 * 32 + 16) * 32, 0), ((0, 0, 0), 0 + (16 * 32 + 16) * 32, 9)])  # glass passes light: no height change
    flags = {tuple(int(v) for v in p): int(f) for p, f in zip(r["remesh"], r["remesh_flags"])}
    assert flags[(0, 0, 0)] & 1 and flags[(-1, 0, 0)] & 1
    m.close()


def test_set_blocks_matches_the_reference_edit_chain_golden():
    """hcm_set_blocks + the mesh kernel against the UNMODIFIED reference's kSetBlock / kFloodSunlight / kSetSunlight / kMesh tasks
    (tests/golden/edit_chain_golden.json, recorded from oracle/_ref by tests/golden/make_edit_golden.py): per batch the number of
    changed blocks, every sunlight height change, the mesh tasks pushed and who pushed them (priority), and the SHA-256 of every
    remeshed chunk's meshes against the digest of the reference's own mesher run after the reference's own edit."""
    import edit_oracle
    from hypercraft_b200 import mesher as M

    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "edit_chain_golden.json")))
    allpos, chunks = edit_oracle.edit_test_world()
    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    m.upload(allpos, np.stack([chunks[tuple(int(c) for c in p)][0] for p in allpos]), np.stack([chunks[tuple(int(c) for c in p)][1] for p in allpos]))
    for rec in golden["batches"]:
        r = m.set_blocks([(tuple(c), i, b) for c, i, b in rec["edits"]])
        assert r["n_changed_blocks"] == rec["n_changed_blocks"], rec["name"]
        got_sun = sorted([[int(v) for v in s["chunk"]], int(s["index"]), int(s["old_height"]), int(s["new_height"])] for s in r["sunlight"])
        assert got_sun == rec["sunlight"], rec["name"]
        want = {tuple(int(v) for v in k.split(",")): e for k, e in rec["remesh"].items()}
        got = {tuple(int(v) for v in p): int(f) for p, f in zip(r["remesh"], r["remesh_flags"])}
        assert set(got) == set(want), rec["name"]
        assert all(bool(got[p] & 1) == bool(want[p]["high"]) for p in want), rec["name"]
        res = m.mesh(r["remesh"])
        for t, p in enumerate(r["remesh"]):
            assert hashlib.sha256(res.serialise(t).tobytes()).hexdigest() == want[tuple(int(v) for v in p)]["mesh_sha256"], (rec["name"], p)
    m.close()


def test_halo_replication_between_contexts(port):
    """hcm_copy_chunks_from (boundary planes from GPU to GPU when the pool is device resident): chunks resident in one context
    become resident in another, device to device, and mesh there like uploaded ones; after hcm_set_blocks on the owner the touched
    chunks are replicated and the holder's meshes follow the edit.  Two contexts on device 0 here (cudaMemcpyPeerAsync between two
    GPUs is the same call; tests/cpp/test_sharded_gpu.cpp covers two devices)."""
    import edit_oracle
    from hypercraft_b200 import mesher as M, registry

    reg = registry.load()
    vpass = reg.props["transparency"][reg.prop_index] == 2
    allpos, chunks = edit_oracle.edit_test_world(31)
    owner = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    owner.upload(allpos, np.stack([chunks[tuple(int(c) for c in p)][0] for p in allpos]), np.stack([chunks[tuple(int(c) for c in p)][1] for p in allpos]))
    holder = M.Mesher(origin=(-1, -2, -2), extent=(4, 5, 5))  # sees the planes x = -1 .. 2
    mine = allpos[allpos[:, 0] >= -1]
    h2d0 = holder.stats()["h2d_bytes"]
    holder.copy_chunks_from(owner, mine)
    tasks = np.array([[x, y, z] for x in (0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)], dtype=np.int16)

    def check(what):
        res = holder.mesh(tasks)
        for t, p in enumerate(tasks):
            got, want = res.serialise(t), port.mesh(chunks, p)
            assert got.size == want.size and (got == want).all(), f"{what}: task {p}\n" + describe_diff(got, want)

    check("replicated world")
    s = holder.stats()
    assert s["p2p_bytes"] == len(mine) * (65536 + 1024) and s["h2d_bytes"] - h2d0 < 65536  # whole chunks from the peer, bookkeeping from the host
    # an edit on the owner: a shaft in chunk (0,0,0) next to the holder's first task plane, a lamp, stone on top of (1,1,0)
    batch = [((0, 0, 0), 3 + (y * 32 + 7) * 32, 0) for y in range(32)] + [((0, 0, 0), 4 + (2 * 32 + 7) * 32, 24), ((1, 1, 0), 9 + (31 * 32 + 9) * 32, 1)]
    r = owner.set_blocks(batch)
    n, sun, _ = edit_oracle.apply_block_edits(chunks, batch, vpass)
    assert r["n_changed_blocks"] == n and len(r["sunlight"]) == len(sun)
    touched = {tuple(c) for c, _, _ in batch} | {tuple(int(v) for v in s_["chunk"]) for s_ in r["sunlight"]}
    holder.copy_chunks_from(owner, np.array(sorted(t for t in touched if t[0] >= -1), dtype=np.int16))
    check("after the owner's edit")
    # errors: not resident in the source, outside the destination's window; nothing changes
    with pytest.raises(M.HcmError) as e:
        owner.copy_chunks_from(holder, np.array([[-2, 0, 0]], dtype=np.int16))  # the holder never had x = -2
    assert e.value.status == -4
    with pytest.raises(M.HcmError) as e:
        holder.copy_chunks_from(owner, np.array([[-2, 0, 0]], dtype=np.int16))
    assert e.value.status == -5
    check("after the failed calls")
    owner.close(), holder.close()
