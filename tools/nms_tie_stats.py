"""Debug aid: tying corners per layer and per-stage times of one 1080p frame (needs a GPU)."""
import sys

import numpy as np
sys.path.insert(0,'.')
import ethzasl_brisk_b200 as bb
from ethzasl_brisk_b200.api import _ptr
ctx = bb.Context(0, timing=True)
img = bb.synthetic_frame(1920,1080,2000)
det = bb.BriskFeatureDetector(60,4,ctx=ctx)
for i in range(3):
    k = det.detect(img)
    r = np.zeros(12,np.int32); ctx._lib.brisk_debug_nms_ties(ctx._h,_ptr(r))
    print(len(k), 'ties per layer', r[:8], {a:round(b,3) for a,b in ctx.last_timing()[0].items() if b>0})
c,lc = det.debug_corners(img); print('corners per layer', lc[:8])
