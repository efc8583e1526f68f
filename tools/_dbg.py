import sys, numpy as np
sys.path.insert(0,'.')
import ethzasl_brisk_b200 as bb
from oracle import restate
g = np.load('tests/golden/brisk_verification.npz')
ctx = bb.Context(0)
for name,img in (('golden0',g['image0']),('syn500',bb.synthetic_frame(500,333,3))):
    kp = restate.agast_detect(img,45,4)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    k1,d1 = ext.compute(img,kp); k2,d2 = restate.describe(img,kp)
    bad = np.where(k1['angle']!=k2['angle'])[0]
    print(name, len(k1), 'angle mismatches', len(bad), 'desc equal', np.array_equal(d1,d2))
    for b in bad[:5]:
        print('   ', b, k1[b], k2[b], 'desc diff bits', int(np.unpackbits(d1[b]^d2[b]).sum()))
    ext0 = bb.BriskDescriptorExtractor(False, True, ctx=ctx)
    a1,e1 = ext0.compute(img,kp); a2,e2 = restate.describe(img,kp,False,True)
    badr = np.where((e1!=e2).any(1))[0]
    print('   rot=False desc mismatching rows', badr[:10], 'of', len(e1))
    # integral check on this image
    print('   integral equal', np.array_equal(ctx.debug_integral(img), restate.integral8(img)))
