"""Shared helpers for the parity tests."""
import numpy as np

from hypercraft_b200 import worldgen


def synth_world(kind, seed, centre=(0, 0, 0), kinds_by_offset=None):
    """27 synthetic chunks around `centre`: (positions int16[27,3], blocks u16[27,32768], heights i8[27,1024], dict)."""
    pos = worldgen.neighbourhood(centre)
    blocks = np.zeros((27, 32768), dtype=np.uint16)
    heights = np.zeros((27, 1024), dtype=np.int8)
    chunks = {}
    for i, p in enumerate(pos):
        k = kind
        if kinds_by_offset is not None:
            k = kinds_by_offset.get(tuple(int(a - b) for a, b in zip(p, centre)), kind)
        blocks[i], heights[i] = worldgen.synth_chunk(k, seed, p)
        chunks[tuple(int(c) for c in p)] = (blocks[i], heights[i])
    return pos, blocks, heights, chunks


def describe_diff(got: np.ndarray, want: np.ndarray) -> str:
    import hcoracle

    try:
        g, w = hcoracle.parse_meshes(got), hcoracle.parse_meshes(want)
    except Exception as e:  # malformed
        return f"sizes {got.size} vs {want.size} ({e})"
    s = [f"meshes {len(g)} vs {len(w)}"]
    for i, (a, b) in enumerate(zip(g, w)):
        s.append(f" mesh{i}: trans {a['transparent']}/{b['transparent']} nv {len(a['vertices'])}/{len(b['vertices'])} "
                 f"ni {len(a['indices'])}/{len(b['indices'])} aabb {a['aabb_min']}{a['aabb_max']} / {b['aabb_min']}{b['aabb_max']}")
        n = min(len(a["vertices"]), len(b["vertices"]))
        bad = np.nonzero(a["vertices"][:n] != b["vertices"][:n])[0]
        if len(bad):
            j = int(bad[0])
            s.append(f"  first vertex diff at {j} ({len(bad)} differ): {int(a['vertices'][j]):016x} vs {int(b['vertices'][j]):016x}")
        n = min(len(a["indices"]), len(b["indices"]))
        bad = np.nonzero(a["indices"][:n] != b["indices"][:n])[0]
        if len(bad):
            s.append(f"  first index diff at {int(bad[0])} ({len(bad)} differ)")
    return "\n".join(s)
