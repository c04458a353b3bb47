"""A second, independent restatement of the narrow phase for the CPU tests: a LITERAL Python replay of the reference's control flow —
GJK::intersect (algorithm/minkowski/gjk/mod.rs:74-113), SimplexProcessor3::reduce_to_closest_feature + check_side
(gjk/simplex/simplex3d.rs:16-89,160-205), EPA3::process with Polytope / Face / remove_or_add_edge (epa/epa3d.rs:16-190),
contact() / point() (:69-94), barycentric_vector (primitive/util.rs:44-60), SupportPoint::from_minkowski (minkowski/mod.rs:33-51)
and the Sphere / Cuboid support functions (primitive/sphere.rs:21-24, cuboid.rs:46-63, util.rs:7-23) — written from the Rust
source with Python lists for the SmallVec / Vec containers (remove, swap, swap_remove, position + remove), one f32 operation at a
time (numpy.float32 scalars: every product and sum is rounded separately, as glam's scalar-math does without FMA).
Only TRANSLATION transforms are supported: Mat4::inverse().transform_vector3 is then the identity and transform_point3 is
`w_axis + p` (glam 0.8: the x, y, z axis products are exact), so nothing of glam's matrix code is restated here.
tests/test_host_cpu.py compares the C++ oracle with this replay bit for bit. TEST INFRASTRUCTURE ONLY."""
import numpy as np

f32 = np.float32
ZERO = (f32(0), f32(0), f32(0))


def v3(x, y, z):
    return (f32(x), f32(y), f32(z))


def add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def scale(a, s):
    return (a[0] * s, a[1] * s, a[2] * s)


def neg(a):
    return (-a[0], -a[1], -a[2])


def dot(a, b):  # glam 0.8 scalar Vec3::dot: (x * x) + (y * y) + (z * z), left to right
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def cross(a, b):  # glam 0.8 scalar Vec3::cross
    return (a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1])


def normalize(a):  # glam 0.8: self * (1.0 / self.length())
    return scale(a, f32(1.0) / np.sqrt(dot(a, a)))


def is_zero(a):
    return a[0] == 0 and a[1] == 0 and a[2] == 0


class Sphere:
    def __init__(self, radius):
        self.radius = f32(radius)

    def support_point(self, direction, pos):  # sphere.rs:21-24
        d = direction  # transform.inverse().transform_vector3(direction) for a translation
        return add(pos, scale(d, self.radius / np.sqrt(dot(d, d))))


class Cuboid:
    def __init__(self, hx, hy, hz):  # half dimensions (Cuboid::new(dim) stores dim / 2)
        hx, hy, hz = f32(hx), f32(hy), f32(hz)
        self.corners = [(hx, hy, hz), (-hx, hy, hz), (-hx, -hy, hz), (hx, -hy, hz), (hx, hy, -hz), (-hx, hy, -hz), (-hx, -hy, -hz), (hx, -hy, -hz)]  # cuboid.rs:46-57

    def support_point(self, direction, pos):  # util.rs:7-23 get_max_point: the FIRST maximum wins (strict >)
        max_p, max_dot = ZERO, f32(-np.inf)
        for v in self.corners:
            d = dot(v, direction)
            if d > max_dot:
                max_p, max_dot = v, d
        return add(pos, max_p)


class SupportPoint:
    def __init__(self, v, sup_a):
        self.v, self.sup_a = v, sup_a


def from_minkowski(left, lpos, right, rpos, direction):  # minkowski/mod.rs:33-51
    l = left.support_point(direction, lpos)
    r = right.support_point(neg(direction), rpos)
    return SupportPoint(sub(l, r), l)


def cross_aba(a, b):
    return cross(cross(a, b), a)


def check_side(abc, ab, ac, ao, simplex, above, ignore_ab):  # simplex3d.rs:160-205 -> new v
    ab_perp = cross(ab, abc)
    if not ignore_ab and dot(ab_perp, ao) > 0:
        del simplex[0]
        return cross_aba(ab, ao)
    ac_perp = cross(abc, ac)
    if dot(ac_perp, ao) > 0:
        del simplex[1]
        return cross_aba(ac, ao)
    if above or dot(abc, ao) > 0:
        return abc
    simplex[0], simplex[1] = simplex[1], simplex[0]
    return neg(abc)


def reduce_to_closest_feature(simplex, v):  # simplex3d.rs:16-89 -> (origin inside, v)
    if len(simplex) == 4:
        a, b, c, d = simplex[3].v, simplex[2].v, simplex[1].v, simplex[0].v
        ao, ab, ac = neg(a), sub(b, a), sub(c, a)
        abc = cross(ab, ac)
        if dot(abc, ao) > 0:
            del simplex[0]
            v = check_side(abc, ab, ac, ao, simplex, True, False)
        else:
            ad = sub(d, a)
            acd = cross(ac, ad)
            if dot(acd, ao) > 0:
                del simplex[2]
                v = check_side(acd, ac, ad, ao, simplex, True, True)
            else:
                adb = cross(ad, ab)
                if dot(adb, ao) > 0:
                    del simplex[1]
                    simplex[0], simplex[1] = simplex[1], simplex[0]
                    v = adb
                else:
                    return True, v
    elif len(simplex) == 3:
        a, b, c = simplex[2].v, simplex[1].v, simplex[0].v
        ao, ab, ac = neg(a), sub(b, a), sub(c, a)
        v = check_side(cross(ab, ac), ab, ac, ao, simplex, False, False)
    elif len(simplex) == 2:
        a, b = simplex[1].v, simplex[0].v
        ao, ab = neg(a), sub(b, a)
        v = cross_aba(ab, ao)
        if is_zero(v):
            v = (f32(0.1), v[1], v[2])
    return False, v


def gjk_intersect(left, lpos, right, rpos, max_iterations=100):  # gjk/mod.rs:74-113 -> simplex or None
    d = sub(rpos, lpos)  # transform_point3(zero) of a translation is its w axis
    if is_zero(d):
        d = v3(1, 1, 1)
    a = from_minkowski(left, lpos, right, rpos, d)
    if dot(a.v, d) <= 0:
        return None
    simplex = [a]
    d = neg(d)
    for _ in range(max_iterations):
        a = from_minkowski(left, lpos, right, rpos, d)
        if dot(a.v, d) <= 0:
            return None
        simplex.append(a)
        inside, d = reduce_to_closest_feature(simplex, d)
        if inside:
            return simplex
    return None


class Face:
    def __init__(self, vertices, a, b, c):  # epa3d.rs:160-170
        ab = sub(vertices[b].v, vertices[a].v)
        ac = sub(vertices[c].v, vertices[a].v)
        self.vertices = (a, b, c)
        self.normal = normalize(cross(ab, ac))
        self.distance = dot(self.normal, vertices[a].v)


def remove_or_add_edge(edges, edge):  # epa3d.rs:183-190
    for i, e in enumerate(edges):
        if edge[0] == e[1] and edge[1] == e[0]:
            del edges[i]
            return
    edges.append(edge)


def barycentric_vector(p, a, b, c):  # primitive/util.rs:44-60
    v0, v1, v2 = sub(b, a), sub(c, a), sub(p, a)
    d00, d01, d11, d20, d21 = dot(v0, v0), dot(v0, v1), dot(v1, v1), dot(v2, v0), dot(v2, v1)
    inv_denom = f32(1.0) / (d00 * d11 - d01 * d01)
    v = (d11 * d20 - d01 * d21) * inv_denom
    w = (d00 * d21 - d01 * d20) * inv_denom
    u = f32(1.0) - v - w
    return u, v, w


def epa_process(simplex, left, lpos, right, rpos, tolerance=f32(0.00001), max_iterations=100):
    """epa3d.rs:16-52 -> (normal, depth, point, iterations, faces) or None"""
    if len(simplex) < 4:
        return None
    vertices = simplex
    faces = [Face(vertices, 3, 2, 1), Face(vertices, 3, 1, 0), Face(vertices, 3, 0, 2), Face(vertices, 2, 0, 1)]  # :172-179
    i = 1
    while True:
        face = faces[0]  # closest_face_to_origin (:111-119): the first minimum; nothing compares below a NaN
        for f in faces[1:]:
            if f.distance < face.distance:
                face = f
        p = from_minkowski(left, lpos, right, rpos, face.normal)
        d = dot(p.v, face.normal)
        if d - face.distance < tolerance or i >= max_iterations:
            u, v, w = barycentric_vector(scale(face.normal, face.distance), vertices[face.vertices[0]].v, vertices[face.vertices[1]].v, vertices[face.vertices[2]].v)
            point = add(add(scale(vertices[face.vertices[0]].sup_a, u), scale(vertices[face.vertices[1]].sup_a, v)), scale(vertices[face.vertices[2]].sup_a, w))
            return face.normal, face.distance, point, i, len(faces)
        # Polytope::add (:121-149)
        edges = []
        k = 0
        while k < len(faces):
            dt = dot(faces[k].normal, sub(p.v, vertices[faces[k].vertices[0]].v))
            if dt > 0:
                removed = faces[k]            # swap_remove(k): the last face moves into slot k
                last = faces.pop()
                if k < len(faces):
                    faces[k] = last
                remove_or_add_edge(edges, (removed.vertices[0], removed.vertices[1]))
                remove_or_add_edge(edges, (removed.vertices[1], removed.vertices[2]))
                remove_or_add_edge(edges, (removed.vertices[2], removed.vertices[0]))
            else:
                k += 1
        n = len(vertices)
        vertices.append(p)
        faces.extend(Face(vertices, n, a, b) for a, b in edges)
        i += 1


def intersection(left, lpos, right, rpos):
    """GJK::intersection(FullResolution) (gjk/mod.rs:317-341) -> None or (normal, depth, point, epa iterations, faces)"""
    simplex = gjk_intersect(left, lpos, right, rpos)
    if simplex is None:
        return None
    return epa_process(simplex, left, lpos, right, rpos)


# ---- GJK::distance (gjk/mod.rs:231-280) with get_closest_point_to_origin (simplex3d.rs:91-113), get_closest_point_on_face
# (:117-149), get_closest_point_on_edge (primitive/util.rs:81-94), Plane::from_points (plane.rs:69-86), Plane x Ray (:117-129) ----
F32_EPSILON = f32(1.1920929e-07)


class ReferencePanic(Exception):
    """an unwrap() on None or a failed assert! in the reference"""


def get_closest_point_on_edge(start, end, point):
    line = sub(end, start)
    line_dir = normalize(line)
    v = sub(point, start)
    d = dot(v, line_dir)
    if d < 0:
        return start
    if d * d > dot(line, line):
        return end
    return add(start, scale(line_dir, d))


def get_closest_point_on_face(a, b, c, normal):
    origin, direction = ZERO, neg(normal)                       # Ray::new(Vec3::zero(), -*normal)
    v0, v1 = sub(b, a), sub(c, a)                               # Plane::from_points(a, b, c).unwrap()
    nrm = cross(v0, v1)
    if is_zero(nrm):
        raise ReferencePanic("Plane::from_points -> None")
    n = normalize(nrm)
    pd = -dot(a, n)
    t = -(pd + dot(origin, n)) / dot(direction, n)              # plane.intersection(&ray)
    if t < 0:                                                   # (a NaN t is not < 0: Some(NaN point), as in Rust)
        return ZERO
    point = add(origin, scale(direction, t))
    u, v, w = barycentric_vector(point, a, b, c)
    in_range = lambda x: x >= 0 and x <= 1
    if not (in_range(u) and in_range(v) and in_range(w)):
        raise ReferencePanic("Simplex processing failed to deduce that this simplex is an edge case")
    return point


def get_closest_point_to_origin(simplex):
    inside, d = reduce_to_closest_feature(simplex, ZERO)
    if inside:
        return d
    if len(simplex) == 1:
        return simplex[0].v
    if len(simplex) == 2:
        return get_closest_point_on_edge(simplex[1].v, simplex[0].v, ZERO)
    return get_closest_point_on_face(simplex[2].v, simplex[1].v, simplex[0].v, d)


def distance(left, lpos, right, rpos, max_iterations=100, distance_tolerance=f32(0.000001)):
    """-> ("some", dp - d0) | ("none",) colliding | ("max_iterations",) | ("panic",)"""
    near_zero = lambda d: abs(d[0] - 0) < F32_EPSILON and abs(d[1] - 0) < F32_EPSILON and abs(d[2] - 0) < F32_EPSILON
    d = sub(rpos, lpos)
    if near_zero(d):
        d = v3(1, 1, 1)
    simplex = [from_minkowski(left, lpos, right, rpos, d), from_minkowski(left, lpos, right, rpos, neg(d))]
    try:
        for _ in range(max_iterations):
            d = get_closest_point_to_origin(simplex)
            if near_zero(d):
                return ("none",)
            d = neg(d)
            p = from_minkowski(left, lpos, right, rpos, d)
            dp = dot(p.v, d)
            d0 = dot(simplex[0].v, d)
            if dp - d0 < distance_tolerance:
                return ("some", dp - d0)
            simplex.append(p)
    except ReferencePanic:
        return ("panic",)
    return ("max_iterations",)
