import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import bam3d_b200 as b3
from bam3d_b200 import broad, scenes
ctx = b3.Context(0)
boxes,_ = scenes.aabb_scene(1<<22)
order = np.lexsort((boxes[:,3], boxes[:,0]))
bs = np.ascontiguousarray(boxes[order])
ctx.set_option("timing", 1)
for it in range(3):
    ctx.read_timing()
    sums, axis = broad.variance3(ctx, bs)
    ms, calls = ctx.read_timing()
    print("variance kernel ms", ms, calls, sums, axis)
