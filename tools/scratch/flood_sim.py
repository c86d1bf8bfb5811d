"""Simulate the level-set sunlight flood on r=6 reference terrain; count per step: live columns, changed columns,
columns whose 5-neighbourhood changed in the previous step, rows with a change."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hcoracle
from hypercraft_b200 import worldgen, registry
R = int(os.environ.get("R", "6"))
meshed, resident = worldgen.load_radius_positions(R)
ref = hcoracle.RefWorld(); ref.generate(resident)
b, h = ref.get_chunks(resident)
reg = registry.load()
ipass = (reg.props["transparency"][reg.prop_index] != 0)
idx = {tuple(int(c) for c in p): i for i, p in enumerate(resident)}
ext = 2 * R + 3
N = ext * 32
passw = np.zeros((N, N, N), dtype=bool)  # [y][z][x]
sun = np.zeros((N, N, N), dtype=bool)
for p, i in idx.items():
    ox, oy, oz = ((c + R + 1) * 32 for c in p)
    blk = b[i].reshape(32, 32, 32)  # [y][z][x]
    passw[oy:oy+32, oz:oz+32, ox:ox+32] = ipass[blk]
    hh = h[i].reshape(32, 32).astype(int)  # [z][x]
    ys = np.arange(32)[:, None, None]
    sun[oy:oy+32, oz:oz+32, ox:ox+32] = ys >= hh[None]
tot = {}
ntask = 0
lim_tab = None
for p in meshed:
    cx, cy, cz = (int(c) + R + 1 for c in p)
    x0, y0, z0 = cx*32-15, cy*32-15, cz*32-15
    P = passw[y0:y0+62, z0:z0+62, x0:x0+62]
    S0 = sun[y0:y0+62, z0:z0+62, x0:x0+62] & P  # hmm: sunlit cells (passable assumed)
    blk_any = b[idx[tuple(int(c) for c in p)]]
    if not S0.any() or P[15:47,15:47,15:47].all() and not (~P).any():
        continue
    G = S0.copy()
    # L1 distance of column from inner square [14,48)
    xs = np.arange(62); d1 = np.maximum(0, np.maximum(14 - xs, xs - 47))
    dist = d1[None, :] + d1[:, None]  # [z][x]
    changed_prev = G.any(axis=0)  # columns changed at step 0
    if not changed_prev.any():
        continue
    ntask += 1
    for step in range(1, 15):
        lim = (14 - step) >= dist
        live = (P & ~G).any(axis=0) & lim
        n = np.zeros_like(G)
        n[1:] |= G[:-1]; n[:-1] |= G[1:]; n[:, 1:] |= G[:, :-1]; n[:, :-1] |= G[:, 1:]; n[:, :, 1:] |= G[:, :, :-1]; n[:, :, :-1] |= G[:, :, 1:]
        add = n & P & ~G
        add &= lim[None]
        chg = add.any(axis=0)
        nb = changed_prev.copy()
        nb[1:] |= changed_prev[:-1]; nb[:-1] |= changed_prev[1:]; nb[:, 1:] |= changed_prev[:, :-1]; nb[:, :-1] |= changed_prev[:, 1:]
        need = live & nb
        rows = changed_prev.any(axis=1)
        rown = rows.copy(); rown[1:] |= rows[:-1]; rown[:-1] |= rows[1:]
        need_row = live & rown[:, None]
        # warp granularity: 32 consecutive columns in linear order (approx: row-major over 62x62)
        def warps(m):
            f = m.reshape(-1); f = np.pad(f, (0, (-len(f)) % 32)).reshape(-1, 32); return int(f.any(axis=1).sum())
        t = tot.setdefault(step, np.zeros(8, dtype=np.int64))
        t += np.array([1, live.sum(), chg.sum(), need.sum(), need_row.sum(), warps(live), warps(need), warps(need_row)])
        G |= add
        changed_prev = chg
        if not chg.any():
            break
print("tasks with sunlight flood:", ntask)
print("step  tasks  live/task  changed/task  need(5-nbhd)/task  need(row)/task  warps: live need need_row (of 121)")
for s in sorted(tot):
    t = tot[s]; k = t[0]
    print(f"{s:3d} {k:6d} {t[1]/k:9.0f} {t[2]/k:9.0f} {t[3]/k:9.0f} {t[4]/k:9.0f}   {t[5]/k:6.1f} {t[6]/k:6.1f} {t[7]/k:6.1f}")
a = sum(t[1] for t in tot.values()); bb = sum(t[3] for t in tot.values()); c = sum(t[4] for t in tot.values())
print("total evaluations: live", a, "need", bb, "need_row", c)
print("warp-slots: live", sum(t[5] for t in tot.values()), "need", sum(t[6] for t in tot.values()), "need_row", sum(t[7] for t in tot.values()))
