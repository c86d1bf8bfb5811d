"""The drop-in boundary compiled against the reference: hypercraft_b200/host/reference_binding/ChunkMeshTask.cpp replaces
the reference's client/src/ChunkMeshTask.cpp and is built with the reference's own Chunk.hpp / ChunkTaskPool.hpp / ChunkMeshTask.inl /
World.hpp / BlockVertex.hpp (oracle/Makefile target `binding` -> oracle/_ref/libhcbind.so), the rest of the client -- task pool, chunk pool,
world -- unmodified.  Mesh tasks are pushed with ChunkTaskPool::Push<kMesh> and run by the reference's own ChunkTaskPool::Run; every
std::vector<hc::client::BlockMesh> that reaches WorldRenderer::PushChunkMesh must equal, byte for byte, what the UNMODIFIED
ChunkTaskRunner<kMesh>::Run pushes for the same chunk (oracle/_ref/libhcref.so)."""
import os

import numpy as np
import pytest


def _run(hcm_library):
    import hcoracle
    from hypercraft_b200 import worldgen

    meshed, resident = worldgen.load_radius_positions(3)
    ref = hcoracle.RefWorld()
    ref.generate(resident)  # the reference's DefaultTerrain
    blocks, heights = ref.get_chunks(resident)
    lo, hi = resident.min(axis=0), resident.max(axis=0)
    bw = hcoracle.BindingWorld(hcm_library, origin=lo, extent=hi - lo + 1, load_radius=6)
    for i, p in enumerate(resident):
        bw.set_chunk(p, blocks[i], heights[i])
    got, n_calls = bw.mesh(meshed)
    assert n_calls == len(meshed) and set(got) == {tuple(int(v) for v in p) for p in meshed}  # every task pushed exactly once
    n_nonempty = 0
    for p in meshed:
        want = ref.mesh(p)  # unmodified ChunkTaskRunner<kMesh>::Run -> PushChunkMesh
        g = got[tuple(int(v) for v in p)]
        assert g.size == want.size and (g == want).all(), f"chunk {p}: PushChunkMesh payload differs from the unmodified runner's"
        n_nonempty += int(want[:4].view(np.uint32)[0]) > 0
    assert n_nonempty > 20
    # a second wave (a remesh at high priority): the chunks are resident, nothing is uploaded again, same payloads
    again, n_calls = bw.mesh(meshed[::3], high=True)
    assert n_calls == len(meshed[::3]) and all((again[k] == got[k]).all() for k in again)
    bw.close()


def test_reference_binding_on_the_emulation(have_ref):
    """CPU tier: the binding's hcm_* calls resolve to the CPU emulation build of the library (tests/cusim)."""
    import hcoracle
    from conftest import build_sim

    if not have_ref or not hcoracle.have_binding():
        pytest.skip("oracle/_ref (reference build) not present")
    _run(build_sim())


@pytest.mark.gpu
def test_reference_binding_on_the_gpu(have_ref):
    import hcoracle
    from hypercraft_b200 import mesher

    if not have_ref or not hcoracle.have_binding():
        pytest.skip("oracle/_ref (reference build) not present")
    if os.environ.get("HCM_TEST_SIM"):
        pytest.skip("covered by test_reference_binding_on_the_emulation")
    _run(mesher.LIB_PATH)
