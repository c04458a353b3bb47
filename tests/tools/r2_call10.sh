#!/bin/bash
mkdir -p gpurun_out
BAM3D_LIB=build/variants/lib_hugetrace.so timeout -s KILL 200 python - <<'PY' 2>&1 | tee gpurun_out/c10_trace.log
import sys
sys.path.insert(0, '.')
import numpy as np
import bam3d_b200 as b3
from tests.pinched import ALL_PINCHED, pinched_scene
ctx = b3.Context(0)
sc = pinched_scene(ALL_PINCHED)
shapes = b3.ShapeSet(ctx, sc["kind"], sc["params"], sc["xform"])
c, st = b3.GJK.new(ctx).intersection_batch(b3.CollisionStrategy.FullResolution, shapes, sc["pairs"])
print(st.tolist())
PY
