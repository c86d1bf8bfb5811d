"""Kernel variants and their on-device qualification (include/hcm.h, "Kernel variants").

derive_kernel and mesh_kernel exist in a baseline form and in a variant form that re-schedules the same arithmetic.  These tests
hold the VARIANT forms to the same bar as the baselines -- the oracle's bytes and the reference's golden digests -- and check the
A/B rounds that decide, on the device, which form a context launches.  (The file runs last: every other GPU test exercises the
baselines, which are what a context starts on.)"""
import hashlib
import json
import os

import numpy as np
import pytest

from util import describe_diff, synth_world

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_golden.json")))
KINDS = ["air", "stone", "checker", "transparent", "mix", "rawmeta", "caves", "flow"]


DERIVE_FORMS = [1, 2, 3]  # quarters + hoist, hoist + L2 prefetch, all three (hcm_kernels.cu: derive_kernel<kForm>)
MESH_FORMS = [1, 2]       # merge height test four rows per round, merge scan two rows per iteration (merge_mask<kVariant>)
FORMS = [(1, 1), (2, 2), (3, 1), (3, 2)]  # (derive, mesh) pairs run through the tests


@pytest.fixture(scope="module", params=FORMS, ids=lambda f: f"derive{f[0]}-mesh{f[1]}")
def variant_mesher(request):
    from hypercraft_b200 import mesher as M

    m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
    m.set_option("derive_variant", request.param[0])
    m.set_option("mesh_variant", request.param[1])
    yield m
    m.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_forced_variants_match_oracle_and_reference_digests(variant_mesher, port, kind, seed):
    m = variant_mesher
    info = m.tune_info()
    assert info["derive"]["state"] == info["mesh"]["state"] == "forced" and info["derive"]["selected"] in DERIVE_FORMS and info["mesh"]["selected"] in MESH_FORMS
    assert info["derive"]["n_variants"] == len(DERIVE_FORMS) and info["mesh"]["n_variants"] == len(MESH_FORMS)
    pos, blocks, heights, chunks = synth_world(kind, seed)
    m.upload(pos, blocks, heights)
    res = m.mesh(np.array([[0, 0, 0]], dtype=np.int16))
    got = res.serialise(0)
    want = port.mesh(chunks, (0, 0, 0))
    assert got.size == want.size and (got == want).all(), describe_diff(got, want)
    assert hashlib.sha256(got.tobytes()).hexdigest() == GOLDEN["cases"][f"{kind}/{seed}"]["mesh_sha256"]  # the reference's own bytes
    if seed < 3:
        light = m.debug_light((0, 0, 0))
        assert hashlib.sha256(light.tobytes()).hexdigest() == GOLDEN["cases"][f"{kind}/{seed}"]["light_sha256"]


def _sphere(radius, port):
    """(meshed, resident, blocks, heights, want digests) of a load sphere: the unmodified reference's terrain and mesher where
    oracle/_ref is present, else the repo's terrain and the port."""
    import hcoracle
    from hypercraft_b200 import worldgen

    meshed, resident = worldgen.load_radius_positions(radius)
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
    return meshed, resident, blocks, heights, want


@pytest.fixture(scope="module")
def sphere5(port):
    return _sphere(5, port)


@pytest.mark.parametrize("form,mesh_form", FORMS[1:])
def test_forced_variants_on_a_terrain_sphere(sphere5, form, mesh_form):
    """Every chunk of an r = 5 load sphere through the variant forms, plain path, re-derive path and pipelined path."""
    import hcoracle
    from hypercraft_b200 import mesher as M

    meshed, resident, blocks, heights, want = sphere5
    m = M.Mesher(origin=(-6, -6, -6), extent=(13, 13, 13), max_chunks=len(resident), max_tasks=len(meshed))
    m.set_option("derive_variant", form)
    m.set_option("mesh_variant", mesh_form)
    m.upload(resident, blocks, heights)
    got = hcoracle.digest_batch(m.mesh(meshed))
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} of {len(meshed)} chunks differ, first: task {meshed[bad[0]]} got {got[bad[0]]} want {want[bad[0]]}"
    m.rederive()
    assert (hcoracle.digest_batch(m.mesh(meshed)) == want).all(), "after a re-derive with the variant"
    assert (hcoracle.digest_batch(m.upload_and_mesh(resident, blocks, heights, meshed)) == want).all(), "pipelined path differs"
    m.close()


def test_ab_rounds_decide_and_never_change_a_byte(sphere5):
    """The qualification itself: with the thresholds lowered to this sphere's size the first two re-derives and the first two mesh
    batches are A/B rounds (baseline, every variant, baseline; hashes compared on the device).  Every batch -- during the rounds and
    after the decision, whichever way it went -- is byte-identical to the CPU mesher; no unit's hash may differ."""
    import hcoracle
    from hypercraft_b200 import mesher as M

    meshed, resident, blocks, heights, want = sphere5
    m = M.Mesher(origin=(-6, -6, -6), extent=(13, 13, 13), max_chunks=len(resident), max_tasks=len(meshed))
    m.upload(resident, blocks, heights)  # below the default thresholds: no round yet
    m.set_option("tune_min_chunks", 64), m.set_option("tune_min_tasks", 64), m.set_option("tune_rounds", 2)
    assert m.tune_info()["mesh"]["state"] == m.tune_info()["derive"]["state"] == "pending"
    launches = []
    for step in range(4):
        k0 = m.stats()["kernel_launches"]
        m.rederive()
        res = m.mesh(meshed)
        launches.append(m.stats()["kernel_launches"] - k0)
        assert (hcoracle.digest_batch(res) == want).all(), f"step {step}"
    info = m.tune_info()
    for k in ("derive", "mesh"):
        assert info[k]["state"] in ("variant adopted", "baseline kept"), info
        assert info[k]["rounds"] == 2 and not any(info[k]["mismatches"]), info
        assert (info[k]["selected"] != 0) == (info[k]["state"] == "variant adopted")
        assert info[k]["baseline_ms"] > 0 and all(t > 0 for t in info[k]["variant_ms"]) and len(info[k]["variant_ms"]) == info[k]["n_variants"]
        if info[k]["selected"]:  # the fastest matching variant, at least 2 % below the baseline
            assert info[k]["variant_ms"][info[k]["selected"] - 1] == min(info[k]["variant_ms"]) < 0.98 * info[k]["baseline_ms"]
    # an A/B step runs the baseline twice and every variant once per stage, with a hash kernel after every run but the last;
    # a decided context launches derive + classify + mesh
    nd, nm = info["derive"]["n_variants"], info["mesh"]["n_variants"]
    assert launches[-1] == 3 and launches[-2] == 3 and launches[1] == (2 * nd + 3) + (2 * nm + 4), launches
    assert (hcoracle.digest_batch(m.upload_and_mesh(resident, blocks, heights, meshed)) == want).all()
    m.close()


def test_a_differing_hash_rejects_the_variant_for_good(sphere5):
    """`tune_fault` (a test knob) corrupts the baseline's hash of one unit before the comparisons: every variant then differs in that
    unit, none may be adopted, the context keeps launching the baseline, and the batches' results are untouched."""
    import hcoracle
    from hypercraft_b200 import mesher as M

    meshed, resident, blocks, heights, want = sphere5
    m = M.Mesher(origin=(-6, -6, -6), extent=(13, 13, 13), max_chunks=len(resident), max_tasks=len(meshed))
    m.upload(resident, blocks, heights)
    m.set_option("tune_min_chunks", 64), m.set_option("tune_min_tasks", 64), m.set_option("tune_fault", 1)
    m.rederive()
    res = m.mesh(meshed)
    assert (hcoracle.digest_batch(res) == want).all()
    info = m.tune_info()
    for k in ("derive", "mesh"):
        assert info[k]["state"] == "pending" and info[k]["rounds"] == 1 and all(n >= 1 for n in info[k]["mismatches"]), info
    m.rederive()
    assert (hcoracle.digest_batch(m.mesh(meshed)) == want).all()
    info = m.tune_info()
    for k in ("derive", "mesh"):
        assert info[k]["state"] == "baseline kept, a variant rejected (bytes differ)" and info[k]["selected"] == 0, info
    k0 = m.stats()["kernel_launches"]
    m.rederive()
    assert (hcoracle.digest_batch(m.mesh(meshed)) == want).all()
    assert m.stats()["kernel_launches"] - k0 == 3  # no further rounds
    m.close()


def test_tuning_can_be_switched_off(sphere5):
    from hypercraft_b200 import mesher as M

    meshed, resident, blocks, heights, _ = sphere5
    m = M.Mesher(origin=(-6, -6, -6), extent=(13, 13, 13), max_chunks=len(resident), max_tasks=len(meshed))
    m.set_option("tune", 0), m.set_option("tune_min_chunks", 1), m.set_option("tune_min_tasks", 1)
    m.upload(resident, blocks, heights)
    m.mesh(meshed)
    info = m.tune_info()
    assert info["derive"]["state"] == info["mesh"]["state"] == "off" and info["mesh"]["selected"] == 0 and info["mesh"]["rounds"] == 0
    with pytest.raises(M.HcmError):
        m.set_option("no_such_option", 1)
    m.close()
