"""Pairs of the C2 scene (scenes.mixed_pairs, default seed) whose EPA polytope pinches — a Cylinder cap against a flat face:
the cap's rim supports are coplanar, so face visibility is rounding noise and the reference's face list grows multiplicatively.
(pair index, faces when the reference's iteration cap ends it): the smallest, a few mid-sized and the two largest of the 44 per
2^20 pairs that pass 1024 faces. Shared by the CPU tests (what the reference returns) and the GPU tests (the device returns it)."""
import numpy as np

PINCHED_PAIRS = ((127779, 1034), (69343, 1060), (75000, 1924), (62559, 6892), (87554, 10382), (67675, 35530), (927736, 121962))
ALL_PINCHED = (62559, 67675, 69343, 75000, 87554, 109200, 115747, 127779, 137860, 145366, 232401, 234041, 248928, 254117, 318621, 366488,
               376816, 482312, 504155, 609907, 611507, 659246, 668634, 727511, 745263, 756401, 774395, 791555, 791917, 793696, 818342, 824480,
               902225, 919543, 927736, 931698, 957085, 961080, 972429, 993791, 1009955, 1021000, 1033052, 1040579)


def pinched_scene(indices):
    from bam3d_b200 import scenes
    parts = [scenes.mixed_pairs(1, s=1.5, start=int(i)) for i in indices]
    sc = dict(parts[0])
    for key in ("kind", "params", "xform"):
        sc[key] = np.concatenate([p[key] for p in parts])
    sc["pairs"] = np.arange(2 * len(indices), dtype=np.uint32).reshape(-1, 2)
    return sc
