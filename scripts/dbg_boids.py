import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from oracle import oracle_py as O
from realtimeparticles_b200 import _abi
from scenarios import make_boids
p = make_boids(M=131072, N=512)
for step in range(50):
    p.step(O.STEP_PHYSICS)
    bad = False
    for f in ("CELL_ID", "PERM", "START_END_CELL", "POS", "VEL", "ACC"):
        a, b = p.get(f)
        if f in ("POS", "VEL", "ACC"):
            a, b = a[:512], b[:512]
            neq = ~((a.view(np.uint32) == b.view(np.uint32)) | (a == b))
        else:
            neq = a != b
        if neq.any():
            idx = np.argwhere(neq)[:5]
            print("step", step, f, "mismatches", int(neq.sum()), "first", idx.tolist())
            for i in idx[:3]:
                print("   oracle", a[tuple(i)] if a.ndim > 1 else a[i[0]], "gpu", b[tuple(i)] if b.ndim > 1 else b[i[0]],
                      "row o", a[i[0]], "row g", b[i[0]])
            bad = True
    if bad:
        break
print("done at step", step)
