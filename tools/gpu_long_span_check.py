"""What the product does on the two fuzz configurations where the reference (AddressSanitizer-clean) and the oracle differ:
spans longer than two rows ("Not reproduced").  Expected: the device flags the batch and the call that
synchronises returns an error -- the frame is refused, not silently different."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases as Cs  # noqa: E402
from softlight_b200 import capi as A, scenes as S  # noqa: E402

CASES = [
    {'seed': 5075456, 'n': 1409, 'width': 302, 'height': 202, 'blend': 3, 'depth_test': 2, 'depth_mask': 0, 'depth_type': 5, 'color_type': 26, 'cull': 2,
     'clip_fraction': 0.7484731382258968, 'clear_depth': -1.0, 'split_draws': 3, 'viewport': (206, 115, 175, 150)},
    {'seed': 385179457, 'n': 947, 'width': 486, 'height': 237, 'blend': 3, 'depth_test': 2, 'depth_mask': 0, 'depth_type': 6, 'color_type': 16, 'cull': 1,
     'clip_fraction': 0.7768553306741711, 'clear_depth': 0.5, 'split_draws': 1},
]
if len(sys.argv) > 1:
    CASES = [eval(open(f).read()) for f in sys.argv[1:]]
RCP = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_frames.npz'))['rcp_table']
for kw in CASES:
    sc = S.random_tris_scene(**kw)
    c1, d1, _ = Cs.render("port", sc, rcp_table=RCP)
    try:
        c0, d0, s0 = Cs.render("cuda", sc, rcp_table=RCP)
        same = np.array_equal(Cs.bits(np.ascontiguousarray(c0)) if c0.dtype.kind == "f" else c0, Cs.bits(np.ascontiguousarray(c1)) if c1.dtype.kind == "f" else c1)
        print("seed", kw["seed"], "rendered without an error; equal to the oracle:", same, flush=True)
    except A.SlError as e:
        print("seed", kw["seed"], "refused:", str(e)[:200], flush=True)
