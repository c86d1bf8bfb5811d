"""Regenerates tests/golden/edit_chain_golden.json from the UNMODIFIED reference edit chain (oracle/_ref/libhcref.so:
ChunkTaskRunner<kSetBlock / kFloodSunlight / kSetSunlight / kMesh>::Run popped and run by the reference's own ChunkTaskPool::Run,
hcoracle.RefEditWorld).  Run where /root/reference exists (`make -C oracle ref` first).  Inputs: the integer-hash terrain world and the
seeded batches of oracle/edit_oracle.py (machine independent).  Recorded per batch: the edits, how many blocks changed, every changed
sunlight height (old -> new), the mesh tasks the reference ran (with the priority they were pushed at), the requests it could never
run (a neighbour chunk does not exist), and the SHA-256 of the meshes the reference's own mesher produced for each remeshed chunk
after the edit."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import edit_oracle  # noqa: E402
import hcoracle  # noqa: E402


def record_batch(rw, keys, batch):
    before = {k: rw.get_chunk(k) for k in keys}
    pushed, high, still, meshes = rw.apply(batch, with_meshes=True)
    after = {k: rw.get_chunk(k) for k in keys}
    sun = [[list(k), int(i), int(before[k][1][i]), int(after[k][1][i])] for k in keys for i in np.nonzero(before[k][1] != after[k][1])[0]]
    remesh = {}
    for i, p in enumerate(pushed):  # the reference may run a position several times while the chain settles: the last run is the final mesh
        key = ",".join(str(int(v)) for v in p)
        e = remesh.setdefault(key, {"high": 0})
        e["high"] |= int(high[i])
        e["mesh_sha256"] = hashlib.sha256(meshes[i].tobytes()).hexdigest()
    return {"edits": [[list(c), int(i), int(b)] for c, i, b in batch],
            "n_changed_blocks": int(sum(int((before[k][0] != after[k][0]).sum()) for k in keys)),
            "sunlight": sorted(sun), "remesh": remesh, "mesh_tasks_run": int(len(pushed)), "still_queued": still}


if __name__ == "__main__":
    allpos, chunks = edit_oracle.edit_test_world()
    rw = hcoracle.RefEditWorld(4)
    for p in allpos:
        rw.set_chunk(p, *chunks[tuple(int(c) for c in p)])
    keys = sorted(chunks)
    out = {"source": "unmodified reference kSetBlock / kFloodSunlight / kSetSunlight / kMesh tasks run by ChunkTaskPool::Run via oracle/_ref "
                     "(hcoracle.RefEditWorld); world and batches: oracle/edit_oracle.py edit_test_world(78), edit_test_batches(seed 21)",
           "batches": []}
    for name, batch in edit_oracle.edit_test_batches(chunks):
        rec = record_batch(rw, keys, batch)
        rec["name"] = name
        out["batches"].append(rec)
        for k in keys:  # the generator reads the current heights: keep the python copy in step with the reference world
            b, h = rw.get_chunk(k)
            chunks[k][0][:] = b
            chunks[k][1][:] = h
    json.dump(out, open(os.path.join(HERE, "edit_chain_golden.json"), "w"), indent=1)
    print("wrote edit_chain_golden.json:", [(b["name"], b["n_changed_blocks"], len(b["sunlight"]), len(b["remesh"])) for b in out["batches"]])
