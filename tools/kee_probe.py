"""K_EE (Dot / RBF) over problem sizes: device time and roofline fraction.  CSPBO_LIB selects an alternative library build.
    python tools/kee_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from csp_bo_b200 import _lib
from tools.sweep import block_line, epack, peaks

fp64, _ = peaks()
for single in (False, True):
    for n_struct in [int(v) for v in os.environ.get('KEE_SIZES', '41,83,208,416,833').split(',')]:
        pe = epack(n_struct, single)
        for fam in (_lib.DOT, _lib.RBF):
            r = block_line(fam, _lib.KEE, pe, pe, True, "%d structures of 192 atoms%s" % (n_struct, ", single species" if single else ""), fp64)
            print(json.dumps({k: (round(r[k], 4) if isinstance(r[k], float) else r[k]) for k in ("family", "what", "ms", "ms_best", "frac")}), flush=True)
        del pe
pe = epack(5352, True, rows=64)
for fam in (_lib.DOT, _lib.RBF):
    r = block_line(fam, _lib.KEE, pe, pe, True, "Si_5352-shaped: 5352 structures of 64 atoms", fp64)
    print(json.dumps({k: (round(r[k], 4) if isinstance(r[k], float) else r[k]) for k in ("family", "what", "ms", "ms_best", "frac")}), flush=True)
