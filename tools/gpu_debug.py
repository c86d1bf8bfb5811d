"""Developer script: quick GPU-vs-oracle comparison with per-stage diagnostics (run under gpurun)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hcoracle  # noqa: E402
from hypercraft_b200 import mesher as M, registry  # noqa: E402
from util import describe_diff, synth_world  # noqa: E402

port = hcoracle.PortOracle(registry.load().blob)
m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
kinds = sys.argv[1:] or ["air", "stone", "caves", "checker", "transparent", "mix", "rawmeta", "flow"]
for kind in kinds:
    pos, blocks, heights, chunks = synth_world(kind, 1)
    m.upload(pos, blocks, heights)
    t0 = time.time()
    got_l = m.debug_light((0, 0, 0))
    want_l = port.light34(chunks, (0, 0, 0))
    bad = np.nonzero(got_l != want_l)[0]
    print(f"[{kind}] light: {len(bad)} bytes differ", flush=True)
    if len(bad):
        for b in bad[:5]:
            x, z, y = b % 34 - 1, (b // 34) % 34 - 1, b // 1156 - 1
            print(f"   at ({x},{y},{z}): got {got_l[b]:#04x} want {want_l[b]:#04x}")
    res = m.mesh(np.array([[0, 0, 0]], dtype=np.int16))
    got, want = res.serialise(0), port.mesh(chunks, (0, 0, 0))
    ok = got.size == want.size and (got == want).all()
    st = m.stats()
    print(f"[{kind}] mesh: {'OK' if ok else 'MISMATCH'}  meshes {int(res.tasks[0]['n_meshes'])} verts {len(res.vertices)} mesh_ms {st['mesh_ms']:.3f}", flush=True)
    if not ok:
        print(describe_diff(got, want))
