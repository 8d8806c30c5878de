"""Do two fused sweeps of different pairs overlap when they are allowed to run at the same time?  (development aid)"""
import sys, time, threading
sys.path.insert(0, ".")
import calibrating_b200 as cb
from calibrating_b200 import synth

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
l, r, _ = synth.rectified_pair(1080, 1920, 128, seed=0)
ms = [cb.StereoSGBM_create(minDisparity=0, numDisparities=128, blockSize=5, P1=600, P2=2400, disp12MaxDiff=1, uniquenessRatio=5,
                           speckleWindowSize=200, speckleRange=2, mode=mode) for _ in range(2)]
for m in ms:
    m.compute(l, r)
print("alone: %.3f ms per aggregation" % ms[0].handle.bench_aggregate(20))
res = [0, 0]
def run(i):
    res[i] = ms[i].handle.bench_aggregate(20)
ts = [threading.Thread(target=run, args=(i,)) for i in range(2)]
t0 = time.time(); [t.start() for t in ts]; [t.join() for t in ts]; wall = time.time() - t0
print("two handles at once: %.3f / %.3f ms per aggregation each, wall %.1f ms for 2 x 21 aggregations -> %.3f ms per pair" % (res[0], res[1], wall * 1e3, wall * 1e3 / 42))
