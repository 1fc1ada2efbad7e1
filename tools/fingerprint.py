"""Byte-level fingerprint of everything the device is given (generated header, integer / float tables,
scaling tables, patterns, tile count, shared memory) for a set of cases.  Host-side refactorings made without
a GPU at hand are checked by running it in two checkouts and comparing the output:

    git worktree add /tmp/wt <validated commit>; (cd /tmp/wt && python tools/fingerprint.py) > a.json
    python tools/fingerprint.py > b.json; diff a.json b.json        # run from the repo root
"""
import sys, hashlib, json
sys.path.insert(0, '.')
import numpy as np
from examples import problems
from examples.cases import lower_case, GOLDEN_CASES, build_golden_problem
from pycollo_b200 import engine as E
cases = [("brachistochrone","lobatto",10,4,None),("brachistochrone","radau",7,5,None),
         ("hypersensitive","radau",5,[4,6,3,5,4],[0.1,0.3,0.15,0.25,0.2]),
         ("cart_pole_swing_up","lobatto",2000,4,None),("cart_pole_swing_up","radau",40,4,None),
         ("double_pendulum","lobatto",6,[4,6,3,5,4,7],None),("free_flying_robot","lobatto",50,5,None),
         ("space_shuttle_reentry","radau",30,6,None),("multiphase_sliding_mass","lobatto",900,4,None),
         ("multiphase_sliding_mass","radau",9,3,None),
         ("delta_iii_launch_vehicle","lobatto",500,4,None),("delta_iii_launch_vehicle","lobatto",12,4,None),
         ("cart_pole_swing_up","lobatto",33333,4,None)]
out = {}
for name, method, K, nodes, sizes in cases:
    try:
        low, meshes, scal = lower_case(getattr(problems, name)(), method, K, nodes, sizes, seed=5)
    except TypeError:
        low, meshes, scal = lower_case(getattr(problems, name)(), method, K, nodes, seed=5)
    h = hashlib.sha256()
    h.update(low.header.encode())
    tb = E.build_tables(low.S, low.layouts)
    for k in sorted(tb):
        h.update(k.encode()); h.update(np.ascontiguousarray(tb[k]).tobytes())
    for a in E.scaling_tables(low.S, low.layouts, *scal):
        h.update(np.ascontiguousarray(a).tobytes())
    for a in low.S.G_structure() + low.S.H_structure():
        h.update(np.ascontiguousarray(a).tobytes())
    h.update(str((low.S.num_tiles, low.S.threads, E.smem_bytes(low.S, low.layouts, low.S.threads))).encode())
    out[f"{name}-{method}-{K}-{nodes}"] = h.hexdigest()[:16]
print(json.dumps(out, indent=0))
