import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hypercraft_b200 import mesher as M
from util import synth_world
m = M.Mesher(origin=(-2, -2, -2), extent=(5, 5, 5))
pos, blocks, heights, chunks = synth_world(sys.argv[1] if len(sys.argv) > 1 else "caves", 1)
m.upload(pos, blocks, heights)
m.sync(); print("upload ok", flush=True)
m.mesh_async(np.array([[0, 0, 0]], dtype=np.int16))
print(m.wait(), flush=True)
