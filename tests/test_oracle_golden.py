"""The CPU oracle (oracle/hc_oracle.cpp) against vectors recorded from the UNMODIFIED reference.

The reference has no tests of its own for this path, so the pin is: digests of the reference's
output on this repo's machine-independent synthetic chunks (tests/golden/synth_golden.json, written by
tests/golden/make_golden.py from oracle/_ref) and, where oracle/_ref is present, a live byte-for-byte comparison on
reference terrain and on the dynamic-mesh generator."""
import hashlib
import json
import os

import numpy as np
import pytest

from util import synth_world

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_golden.json")))
MIXED = [("stone", {(1, 0, 0): "air", (0, 1, 0): "mix", (0, 0, -1): "flow"}),
         ("air", {(0, -1, 0): "stone", (1, 0, 0): "mix"}),
         ("flow", {(1, 0, 0): "mix", (-1, 0, 0): "flow", (0, 0, 1): "stone", (0, 0, -1): "rawmeta"}),
         ("caves", {(0, 1, 0): "air", (0, -1, 0): "stone", (1, 1, 0): "mix"})]


def _check(port, chunks, want):
    import hcoracle

    buf = port.mesh(chunks, (0, 0, 0))
    ms = hcoracle.parse_meshes(buf)
    assert [[m["transparent"], len(m["vertices"]), len(m["indices"])] for m in ms] == want["meshes"]
    assert hashlib.sha256(buf.tobytes()).hexdigest() == want["mesh_sha256"]
    assert hashlib.sha256(port.light34(chunks, (0, 0, 0)).tobytes()).hexdigest() == want["light_sha256"]


@pytest.mark.parametrize("case", sorted(GOLDEN["cases"]))
def test_port_matches_reference_golden(port, case):
    kind, seed = case.split("/")
    _, _, _, chunks = synth_world(kind, int(seed))
    _check(port, chunks, GOLDEN["cases"][case])


@pytest.mark.parametrize("i", range(len(MIXED)))
def test_port_matches_reference_golden_mixed(port, i):
    kind, by_off = MIXED[i]
    _, _, _, chunks = synth_world(kind, 10 + i, kinds_by_offset=by_off)
    _check(port, chunks, GOLDEN["mixed"][i])


TERRAIN = [(12314524, (0, 0, 0)), (12314524, (0, 1, 2)), (12314524, (2, -1, 1)), (7, (-2, 0, -1)), (7, (0, 0, 1)), (7, (-1, -1, -2))]


@pytest.mark.parametrize("i", range(len(TERRAIN)))
def test_port_matches_reference_golden_terrain(port, i):
    """Surface chunks of the repo's integer-hash height-field terrain (oceans, plants, lamps; sunlight heights below the
    chunk top): the reference's digests for six centres."""
    import hcoracle
    from hypercraft_b200 import worldgen

    seed, centre = TERRAIN[i]
    chunks = {tuple(int(c) for c in p): worldgen.synth_terrain_chunk(seed, p) for p in worldgen.neighbourhood(centre)}
    want = GOLDEN["terrain"][i]
    buf = port.mesh(chunks, centre)
    assert [[m["transparent"], len(m["vertices"]), len(m["indices"])] for m in hcoracle.parse_meshes(buf)] == want["meshes"]
    assert hashlib.sha256(buf.tobytes()).hexdigest() == want["mesh_sha256"]
    assert hashlib.sha256(port.light34(chunks, centre).tobytes()).hexdigest() == want["light_sha256"]


def test_known_structure_checkerboard(port):
    """3-D checkerboard: 98 304 quads = 393 216 vertices, cut every 65 532 vertices, first index quad of every
    continuation mesh carries the stale cur_vertex 65532..65535 (BlockMeshAlgo.hpp:397-403 then :440-459)."""
    import hcoracle

    _, _, _, chunks = synth_world("checker", 1)
    ms = hcoracle.parse_meshes(port.mesh(chunks, (0, 0, 0)))
    # the halo is a checkerboard of the same parity, so only in-chunk faces towards air are emitted
    assert all(len(m["vertices"]) == 65532 for m in ms[:-1])
    for m in ms[1:]:
        assert sorted(set(int(v) for v in m["indices"][:6])) == [65532, 65533, 65534, 65535]
        assert int(m["indices"][6:12].min()) == 4
    assert list(ms[0]["indices"][:6]) in ([0, 1, 2, 0, 2, 3], [1, 2, 3, 0, 1, 3])


def test_empty_chunks_give_zero_meshes(port):
    for kind in ("air", "stone"):
        _, _, _, chunks = synth_world(kind, 1)
        buf = port.mesh(chunks, (0, 0, 0))
        assert buf.size == 4 and int(buf.view(np.uint32)[0]) == 0


def test_port_matches_live_reference_on_terrain(port, have_ref):
    if not have_ref:
        pytest.skip("oracle/_ref/libhcref.so not present")
    import hcoracle
    from hypercraft_b200 import worldgen

    meshed, resident = worldgen.load_radius_positions(2)
    ref = hcoracle.RefWorld()
    ref.generate(resident)
    b, h = ref.get_chunks(resident)
    chunks = {tuple(int(c) for c in p): (b[i], h[i]) for i, p in enumerate(resident)}
    nonempty = 0
    for p in meshed:
        a, c = ref.mesh(p), port.mesh(chunks, p)
        assert a.size == c.size and (a == c).all(), f"terrain chunk {p}"
        assert (ref.light34(p) == port.light34(chunks, p)).all()
        nonempty += a.size > 4
    assert nonempty >= 5


def test_registry_blob_matches_live_reference(have_ref, registry_blob):
    if not have_ref:
        pytest.skip("oracle/_ref/libhcref.so not present")
    import hcoracle

    live = hcoracle.RefWorld().registry_blob()
    assert live.size == registry_blob.size and (live == registry_blob).all()


def test_native_liquid_generator_matches_reference_function_pointer(port, have_ref):
    """flow_mesh_gen (Liquid.hpp:171-282) is a function pointer in the reference registry; the native restatement is
    pinned by calling the real pointer on random neighbourhoods, raw out-of-table meta values included."""
    if not have_ref:
        pytest.skip("oracle/_ref/libhcref.so not present")
    import hcoracle

    ref = hcoracle.RefWorld()
    rng = np.random.default_rng(5)
    ids = np.array([0, 1, 14, 15, 15, 15, 9, 16, 30, 200], dtype=np.uint16)
    for it in range(4000):
        blk = 15 | (int(rng.integers(0, 256)) << 8)
        meta = rng.integers(0, 256 if it % 2 else 9, 8).astype(np.uint16)
        nb = (ids[rng.integers(0, len(ids), 8)] | (meta << 8)).astype(np.uint16)
        ra, pa = ref.dynamic_mesh_eval(blk, nb), port.dynamic_mesh_eval(blk, nb)
        assert ra[0] == pa[0] and (ra[1] == pa[1]).all() and (ra[2] == pa[2]).all() and ra[3] == pa[3]


# ---- the edit chain: oracle/edit_oracle.py pinned to the reference's own tasks ---------------------------
EDIT_GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "edit_chain_golden.json")))


def _vertical_pass():
    from hypercraft_b200 import registry

    reg = registry.load()
    return reg.props["transparency"][reg.prop_index] == 2


def test_edit_oracle_matches_reference_edit_chain_golden(port):
    """The restatement of the kSetBlock -> kFloodSunlight -> kSetSunlight chain against what the UNMODIFIED reference tasks did on the
    same world and batches (tests/golden/edit_chain_golden.json, recorded by tests/golden/make_edit_golden.py): blocks changed,
    every sunlight height changed, the mesh tasks run and their priority, and -- through the CPU mesher port on the edited world -- the
    bytes of every remeshed chunk against the digests of the reference's own mesher."""
    import edit_oracle

    vpass = _vertical_pass()
    _, chunks = edit_oracle.edit_test_world()
    for rec in EDIT_GOLDEN["batches"]:
        batch = [(tuple(c), i, b) for c, i, b in rec["edits"]]
        n, sun, remesh, flags = edit_oracle.apply_block_edits(chunks, batch, vpass, with_flags=True)
        assert n == rec["n_changed_blocks"], rec["name"]
        assert sorted([list(k[0]), k[1], v[0], v[1]] for k, v in sun.items()) == rec["sunlight"], rec["name"]
        meshable = {p for p in remesh if all(abs(v) <= 1 for v in p)}  # the 27 chunks around must exist (5^3 world)
        want = {tuple(int(v) for v in k.split(",")): e for k, e in rec["remesh"].items()}
        assert meshable == set(want), rec["name"]
        assert len(remesh) - len(meshable) <= rec["still_queued"]  # requests the reference keeps queued for ever
        for p, e in want.items():
            assert bool(flags[p] & 1) == bool(e["high"]), (rec["name"], p)
            assert hashlib.sha256(port.mesh(chunks, p).tobytes()).hexdigest() == e["mesh_sha256"], (rec["name"], p)


def test_edit_oracle_matches_reference_edit_chain_live(have_ref):
    """The same comparison live, on other seeds, where oracle/_ref is present: block and height state of every chunk after every batch."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    import edit_oracle
    import hcoracle

    vpass = _vertical_pass()
    for world_seed, batch_seed in [(5, 1), (1234, 99)]:
        allpos, chunks = edit_oracle.edit_test_world(world_seed)
        rw = hcoracle.RefEditWorld(4)
        for p in allpos:
            rw.set_chunk(p, *chunks[tuple(int(c) for c in p)])
        for name, batch in edit_oracle.edit_test_batches(chunks, seed=batch_seed, n_random=6):
            n, sun, remesh, flags = edit_oracle.apply_block_edits(chunks, batch, vpass, with_flags=True)
            pushed, high, _, _ = rw.apply(batch)
            for k, (b, h) in chunks.items():
                rb, rh = rw.get_chunk(k)
                assert (rb == b).all() and (rh == h).all(), (world_seed, name, k)
            ran = {}
            for p, hi in zip(pushed, high):
                ran[tuple(int(v) for v in p)] = ran.get(tuple(int(v) for v in p), 0) | int(hi)
            assert set(ran) == {p for p in remesh if all(abs(v) <= 1 for v in p)}, (world_seed, name)
            assert all(bool(flags[p] & 1) == bool(hi) for p, hi in ran.items()), (world_seed, name)
