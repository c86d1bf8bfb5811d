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
