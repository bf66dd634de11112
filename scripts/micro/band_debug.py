import faulthandler, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "optical-flow-python_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, synth
from optical_flow import _lib, load_of_method
from optical_flow.utils.derivatives import partial_deriv
from optical_flow.rowband import RowBand, arena_bytes_for
os.environ["B200FLOW_BAND_MIN_PIXELS"] = "1024"
H, W = 136, 200
im1, im2, flow = synth.gray_pair(H, W, seed=5)
images = np.stack([im1, im2], axis=2); uv = 0.8 * flow
def solve(tag):
    t0 = time.time()
    ope = load_of_method("classic++"); ope.images = images
    It, Ix, Iy = partial_deriv(images, uv, ope.interpolation_method, ope.deriv_filter, ope.blend)
    print(tag, "partial_deriv done %.3f" % (time.time() - t0), flush=True)
    A = ope.flow_operator(uv, np.zeros_like(uv), It, Ix, Iy)[0]
    b = A.b
    print(tag, "A.b done %.3f" % (time.time() - t0), flush=True)
    x = ope._solve_linear_system(A, b, uv.shape)
    print(tag, "solve done %.3f" % (time.time() - t0), ope.last_stats, flush=True)
    return x
solve("main")
bar = threading.Barrier(2); exports = [None, None]
def worker(rank):
    ctx = _lib.default_context(0)
    rb = RowBand(ctx, rank, 2, arena_bytes_for(H, W), same_device=True)
    exports[rank] = rb.export(); bar.wait()
    rb.connect(exports, same_process=True); bar.wait()
    solve("rank%d" % rank)
    bar.wait()
    try:
        rb.close()
    except Exception as e:
        print("close:", e)
faulthandler.dump_traceback_later(4, repeat=False)
th = [threading.Thread(target=worker, args=(r,)) for r in range(2)]
[t.start() for t in th]; [t.join() for t in th]
