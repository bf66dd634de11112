import sys, time, numpy as np
sys.path.insert(0, 'optical-flow-python_b200'); sys.path.insert(0, 'oracle')
import optical_flow as of
from optical_flow import _lib
d = np.load('tests/golden/rubberwhale_10_11.npz'); g = np.load('tests/golden/rubberwhale_full.npz')
ctx = _lib.default_context(); ctx.set_timing(True)
for rtol in (1e-8, 1e-9):
    for rep in range(2):
        t0 = time.time()
        ope = of.load_of_method('classic+nl-fast'); ope.display = False; ope.exact_rtol = rtol
        from optical_flow.interface import _rgb2gray, _rgb2lab
        im1 = d['im1'].astype(float); im2 = d['im2'].astype(float)
        ope.images = np.stack([_rgb2gray(im1), _rgb2gray(im2)], 2); ope.color_images = _rgb2lab(im1, True)
        uv = ope.compute_flow(np.zeros((388,584,2)))
        dt = time.time()-t0
    aae, std, aepe = of.flow_angular_error(d['tu'], d['tv'], uv[:,:,0], uv[:,:,1], 0)
    print('rtol', rtol, 'wall %.3f s' % dt, 'max|d| %.3e' % np.abs(uv-g['uv']).max(), 'AAE %.6f (ref %.6f) AEPE %.6f (ref %.6f)' % (aae, g['aae'], aepe, g['aepe']))
    print('  stats', ope.last_stats)
ims1 = np.stack([d['im1']]*8); ims2 = np.stack([d['im2']]*8)
for rep in range(2):
    t0 = time.time(); uvb, st = of.estimate_flow_batch(ims1, ims2, 'classic+nl-fast', return_stats=True); dt = time.time()-t0
print('batch 8: wall %.3f s -> %.2f pairs/s' % (dt, 8/dt), 'max|d| vs ref %.3e' % np.abs(uvb[0]-g['uv']).max(), st)
