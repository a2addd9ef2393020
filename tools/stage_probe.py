import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import sweep
from csp_bo_b200 import _lib
fp64, hbm = sweep.peaks()
pf = sweep.fpack(6667, False); pe = sweep.epack(833, False)
out = {}
for fam in (_lib.DOT, _lib.RBF):
    for blk, a, b, sym in ((_lib.KEE, pe, pe, True), (_lib.KEF, pe, pf, False)):
        r = sweep.block_line(fam, blk, a, b, sym, "", fp64)
        out["%s_%s" % (r["block"], r["family"])] = round(r["ms"], 3)
print(json.dumps(dict(ee=os.environ.get("CSPBO_STAGE_KB_EE"), ef=os.environ.get("CSPBO_STAGE_KB_EF"), **out)), flush=True)
