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
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "test_dispatcher")
    src = exe + ".cpp"
    lib = os.path.join(root, "hypercraft_b200", "csrc")
    subprocess.run(["g++", "-std=c++20", "-O2", "-pthread", "-o", exe, src, f"-L{lib}", "-lhcm", "-ldl", f"-Wl,-rpath,{lib}"], check=True)
    import hcoracle

    hcoracle.build(ref=False, port=True)
    r = subprocess.run([exe, os.path.join(root, "hypercraft_b200", "data", "registry.bin"), hcoracle.PORT_SO], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1].startswith("PASS"), r.stdout + r.stderr


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
    from hypercraft_b200 import edits, mesher as M, registry, worldgen

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
    from hypercraft_b200 import edits, mesher as M, registry, worldgen

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
