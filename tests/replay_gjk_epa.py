"""A second, independent restatement of the narrow phase for the CPU tests: a LITERAL Python replay of the reference's control flow —
GJK::intersect (algorithm/minkowski/gjk/mod.rs:74-113), SimplexProcessor3::reduce_to_closest_feature + check_side
(gjk/simplex/simplex3d.rs:16-89,160-205), EPA3::process with Polytope / Face / remove_or_add_edge (epa/epa3d.rs:16-190),
contact() / point() (:69-94), barycentric_vector (primitive/util.rs:44-60), SupportPoint::from_minkowski (minkowski/mod.rs:33-51)
and the Particle / Quad / Sphere / Cuboid / Cylinder / Capsule support functions (primitive3.rs:116-131, quad.rs:45-59,
sphere.rs:21-24, cuboid.rs:46-63, util.rs:7-23, cylinder.rs:38-61, capsule.rs:41-48) — written from the Rust
source with Python lists for the SmallVec / Vec containers (remove, swap, swap_remove, position + remove), one f32 operation at a
time (numpy.float32 scalars: every product and sum is rounded separately, as glam's scalar-math does without FMA).
Transforms are full Mat4s: glam 0.8's scalar Mat4::inverse (cofactor expansion with its sign vectors), transform_vector3 and
transform_point3 are restated below as well — they have to be, even for translations: the inverse of a translation has -0.0 in
some off-diagonal entries, so a direction component of exactly +-0 can change its sign on the way through, and Cylinder / Capsule
supports branch on that sign (an axis-aligned EPA face normal is enough to get there). The control flow is what this replay
restates independently; the glam formulas are the same assumption the oracle makes (glam is not vendored in the reference).
tests/test_host_cpu.py compares the C++ oracle with this replay bit for bit. TEST INFRASTRUCTURE ONLY."""
import numpy as np

f32 = np.float32
ZERO = (f32(0), f32(0), f32(0))
F32_EPS = f32(1.1920929e-07)  # f32::EPSILON


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


class Mat4:
    """glam 0.8 Mat4, column-major: cols[c] = (x, y, z, w) of column c."""

    def __init__(self, cols16):
        m = [f32(x) for x in cols16]
        self.c = [tuple(m[4 * k:4 * k + 4]) for k in range(4)]
        self._inv = None

    @staticmethod
    def from_translation(t):
        return Mat4([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, t[0], t[1], t[2], 1])

    def transform_point3(self, p):   # r = x_axis * p.x; r = y_axis * p.y + r; r = z_axis * p.z + r; r = w_axis + r
        x, y, z, w = (c[:3] for c in self.c)
        r = scale(x, p[0])
        r = add(scale(y, p[1]), r)
        r = add(scale(z, p[2]), r)
        return add(w, r)

    def transform_vector3(self, v):
        x, y, z = (c[:3] for c in self.c[:3])
        r = scale(x, v[0])
        r = add(scale(y, v[1]), r)
        return add(scale(z, v[2]), r)

    def inverse(self):
        if self._inv is not None:
            return self._inv
        (m00, m01, m02, m03), (m10, m11, m12, m13), (m20, m21, m22, m23), (m30, m31, m32, m33) = self.c
        coef00 = m22 * m33 - m32 * m23; coef02 = m12 * m33 - m32 * m13; coef03 = m12 * m23 - m22 * m13
        coef04 = m21 * m33 - m31 * m23; coef06 = m11 * m33 - m31 * m13; coef07 = m11 * m23 - m21 * m13
        coef08 = m21 * m32 - m31 * m22; coef10 = m11 * m32 - m31 * m12; coef11 = m11 * m22 - m21 * m12
        coef12 = m20 * m33 - m30 * m23; coef14 = m10 * m33 - m30 * m13; coef15 = m10 * m23 - m20 * m13
        coef16 = m20 * m32 - m30 * m22; coef18 = m10 * m32 - m30 * m12; coef19 = m10 * m22 - m20 * m12
        coef20 = m20 * m31 - m30 * m21; coef22 = m10 * m31 - m30 * m11; coef23 = m10 * m21 - m20 * m11
        fac = [(coef00, coef00, coef02, coef03), (coef04, coef04, coef06, coef07), (coef08, coef08, coef10, coef11),
               (coef12, coef12, coef14, coef15), (coef16, coef16, coef18, coef19), (coef20, coef20, coef22, coef23)]
        vec = [(m10, m00, m00, m00), (m11, m01, m01, m01), (m12, m02, m02, m02), (m13, m03, m03, m03)]
        comb = lambda a, fa, b, fb, c, fc: tuple(a[k] * fa[k] - b[k] * fb[k] + c[k] * fc[k] for k in range(4))
        inv0 = comb(vec[1], fac[0], vec[2], fac[1], vec[3], fac[2])
        inv1 = comb(vec[0], fac[0], vec[2], fac[3], vec[3], fac[4])
        inv2 = comb(vec[0], fac[1], vec[1], fac[3], vec[3], fac[5])
        inv3 = comb(vec[0], fac[2], vec[1], fac[4], vec[2], fac[5])
        sa, sb = (f32(1), f32(-1), f32(1), f32(-1)), (f32(-1), f32(1), f32(-1), f32(1))
        cols = [tuple(inv0[k] * sa[k] for k in range(4)), tuple(inv1[k] * sb[k] for k in range(4)),
                tuple(inv2[k] * sa[k] for k in range(4)), tuple(inv3[k] * sb[k] for k in range(4))]
        d = [self.c[0][k] * cols[k][0] for k in range(4)]
        rcp_det = f32(1.0) / (((d[0] + d[1]) + d[2]) + d[3])
        out = Mat4([0] * 16)
        out.c = [tuple(x * rcp_det for x in col) for col in cols]
        self._inv = out
        return out


class Sphere:
    def __init__(self, radius):
        self.radius = f32(radius)

    def support_point(self, direction, transform):  # sphere.rs:21-24
        d = transform.inverse().transform_vector3(direction)
        return transform.transform_point3(scale(d, self.radius / np.sqrt(dot(d, d))))


class Cuboid:
    def __init__(self, hx, hy, hz):  # half dimensions (Cuboid::new(dim) stores dim / 2)
        hx, hy, hz = f32(hx), f32(hy), f32(hz)
        self.corners = [(hx, hy, hz), (-hx, hy, hz), (-hx, -hy, hz), (hx, -hy, hz), (hx, hy, -hz), (-hx, hy, -hz), (-hx, -hy, -hz), (hx, -hy, -hz)]  # cuboid.rs:46-57

    def support_point(self, direction, transform):  # util.rs:7-23 get_max_point: the FIRST maximum wins (strict >)
        direction = transform.inverse().transform_vector3(direction)
        max_p, max_dot = ZERO, f32(-np.inf)
        for v in self.corners:
            d = dot(v, direction)
            if d > max_dot:
                max_p, max_dot = v, d
        return transform.transform_point3(max_p)


class Quad:
    def __init__(self, hx, hy):  # half dimensions
        hx, hy = f32(hx), f32(hy)
        self.corners = [(hx, hy, f32(0)), (-hx, hy, f32(0)), (-hx, -hy, f32(0)), (hx, -hy, f32(0))]  # quad.rs:45-52

    support_point = Cuboid.support_point  # quad.rs:55-59: get_max_point over its corners


class ConvexPolyhedron:
    def __init__(self, vertices):  # ConvexPolyhedron::new (VertexOnly mode): brute-force support, polyhedron.rs:135-150, 342-355
        self.corners = [v3(*v) for v in vertices]

    support_point = Cuboid.support_point  # the same fold: first maximum over the vertices in their stored order


class ConvexPolyhedronWithFaces:
    """ConvexPolyhedron::new_with_faces (HalfEdge mode): build_half_edges (polyhedron.rs:246-340) and hill_climb_support_point
    (:153-183). Only the links the climb follows are kept: per vertex its first outgoing edge, per edge target / twin / next."""

    def __init__(self, vertices, faces):
        self.pos = [v3(*v) for v in vertices]
        n = len(self.pos)
        self.vertex_edge, vertex_ready = [0] * n, [False] * n
        self.target, self.twin, self.next_edge = [], [], []
        edge_map = {}
        for a, b, c in faces:
            fv = (int(a), int(b), int(c))
            face_edges = [0, 0, 0]
            for j in range(3):
                i = 2 if j == 0 else j - 1
                v0, v1 = fv[i], fv[j]
                if (v0, v1) in edge_map:
                    e01 = edge_map[(v0, v1)]
                    e10 = self.twin[e01]
                else:
                    e01, e10 = len(self.target), len(self.target) + 1
                    self.target += [v1, v0]
                    self.twin += [e10, e01]
                    self.next_edge += [0, 0]
                edge_map[(v0, v1)] = e01
                edge_map[(v1, v0)] = e10
                if not vertex_ready[v0]:
                    self.vertex_edge[v0] = e01
                    vertex_ready[v0] = True
                face_edges[i] = e01
            for j in range(3):
                i = 2 if j == 0 else j - 1
                self.next_edge[face_edges[i]] = face_edges[j]

    def support_point(self, direction, transform):
        d = transform.inverse().transform_vector3(direction)
        best_index = 0
        best_dot = dot(self.pos[best_index], d)
        while True:
            previous_best_dot = best_dot
            edge_index = self.vertex_edge[best_index]
            start_edge_index = edge_index
            while True:
                vertex_index = self.target[edge_index]
                dt = dot(self.pos[vertex_index], d)
                if dt > best_dot:
                    best_index, best_dot = vertex_index, dt
                edge_index = self.next_edge[self.twin[edge_index]]
                if start_edge_index == edge_index:
                    break
            if abs(best_dot - previous_best_dot) < F32_EPS:
                break
        return transform.transform_point3(self.pos[best_index])


class Particle:
    def support_point(self, direction, transform):  # primitive3.rs:119 / particle.rs: the transformed origin
        return transform.transform_point3(ZERO)


class Cylinder:
    def __init__(self, half_height, radius):
        self.half_height, self.radius = f32(half_height), f32(radius)

    def support_point(self, direction, transform):  # cylinder.rs:38-61 (Y-axis aligned)
        result = transform.inverse().transform_vector3(direction)
        negative = bool(np.signbit(result[1]))               # f32::is_sign_negative (true for -0.0 and negative NaN)
        result = (result[0], f32(0), result[2])
        if abs(dot(result, result) - 0) < F32_EPS:
            result = ZERO
        else:
            result = normalize(result)
            if abs(result[1] - 0) < F32_EPS and abs(result[0] - 0) < F32_EPS and abs(result[2] - 0) < F32_EPS:
                result = ZERO
            else:
                result = scale(result, self.radius)
        result = (result[0], -self.half_height if negative else self.half_height, result[2])
        return transform.transform_point3(result)


class Capsule:
    def __init__(self, half_height, radius):
        self.half_height, self.radius = f32(half_height), f32(radius)

    def support_point(self, direction, transform):  # capsule.rs:41-48
        direction = transform.inverse().transform_vector3(direction)
        sy = direction[1]
        sign = f32(np.nan) if sy != sy else (f32(-1.0) if np.signbit(sy) else f32(1.0))   # f32::signum: +-1 by sign bit, NaN for NaN
        result = (f32(0), sign * self.half_height, f32(0))
        return transform.transform_point3(add(result, scale(direction, self.radius / np.sqrt(dot(direction, direction)))))


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
    left_pos = lpos.transform_point3(ZERO)
    right_pos = rpos.transform_point3(ZERO)
    d = sub(right_pos, left_pos)
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
    right_pos = rpos.transform_point3(ZERO)
    left_pos = lpos.transform_point3(ZERO)
    d = sub(right_pos, left_pos)
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


# ---- GJK::intersection_time_of_impact (gjk/mod.rs:130-217), without the contact point (that goes through glam's quaternion
# decomposition in translation_interpolate, traits.rs:180-190, which is not restated here) ----
def time_of_impact(left, l0, l1, right, r0, r1, continuous_tolerance=f32(0.000001), cap=100):
    """-> ("some", lambda, normal, depth) | ("none",) | ("panic",) | ("cap",): the reference's while loop has no bound (the device and
    the oracle stop after `cap` trips and report STATUS_MAX_ITER)"""
    left_lin_vel = sub(l1.transform_point3(ZERO), l0.transform_point3(ZERO))
    right_lin_vel = sub(r1.transform_point3(ZERO), r0.transform_point3(ZERO))
    ray = sub(right_lin_vel, left_lin_vel)
    lam = f32(0)
    normal, ray_origin = ZERO, ZERO
    simplex = []
    p = from_minkowski(left, l0, right, r0, neg(ray))
    v = p.v
    trips = 0
    try:
        while dot(v, v) > continuous_tolerance:
            if trips >= cap:
                return ("cap",)
            trips += 1
            p = from_minkowski(left, l0, right, r0, neg(v))
            vp, vr = dot(v, p.v), dot(v, ray)
            if vp > lam * vr:
                if vr > 0:
                    lam = vp / vr
                    if lam > 1:
                        return ("none",)
                    ray_origin = scale(ray, lam)
                    simplex.clear()
                    normal = neg(v)
                else:
                    return ("none",)
            simplex.append(SupportPoint(sub(p.v, ray_origin), p.sup_a))
            v = get_closest_point_to_origin(simplex)
    except ReferencePanic:
        return ("panic",)
    if dot(v, v) <= continuous_tolerance:
        return ("some", lam, neg(normalize(normal)), np.sqrt(dot(v, v)))
    return ("none",)


# ---- compound shapes (gjk/mod.rs:362-456, 478-523): every primitive x primitive pair under mul_mat4-composed transforms, then Rust's
# Iterator::max_by (the LAST maximum) with partial_cmp().unwrap() / min_by (the FIRST minimum) with unwrap_or(Equal) ----
def mul_mat4(a, b):
    """glam 0.8 Mat4::mul_mat4 (a * b): column j = a.x_axis * b[j].x; a.y_axis * b[j].y + r; a.z_axis * b[j].z + r; a.w_axis * b[j].w + r"""
    cols = []
    for v in b.c:
        r = tuple(a.c[0][k] * v[0] for k in range(4))
        for ax in (1, 2, 3):
            r = tuple(a.c[ax][k] * v[ax] + r[k] for k in range(4))
        cols.extend(r)
    return Mat4(cols)


def intersection_complex(left, lt, right, rt, collision_only=False):
    """left / right: lists of (primitive, local Mat4). -> None | "hit" (CollisionOnly) | (normal, depth, point) | raises ReferencePanic"""
    contacts = []
    for lp, ll in left:
        ltc = mul_mat4(lt, ll)
        for rp, rl in right:
            rtc = mul_mat4(rt, rl)
            if collision_only:
                if gjk_intersect(lp, ltc, rp, rtc) is not None:
                    return "hit"
                continue
            c = intersection(lp, ltc, rp, rtc)
            if c is not None:
                contacts.append(c[:3])
    best = None
    for c in contacts:                                   # max_by: a later element that compares Equal or Greater replaces the current one
        if best is not None and (c[1] != c[1] or best[1] != best[1]):
            raise ReferencePanic("partial_cmp(NaN).unwrap()")
        if best is None or not (c[1] < best[1]):
            best = c
    return best


def distance_complex(left, lt, right, rt):
    """-> None as soon as one primitive pair returns None, else the f32::min of all (NaN-ignoring, as f32::min is)"""
    best = None
    for lp, ll in left:
        ltc = mul_mat4(lt, ll)
        for rp, rl in right:
            rtc = mul_mat4(rt, rl)
            r = distance(lp, ltc, rp, rtc)
            if r[0] == "panic":
                raise ReferencePanic("distance")
            if r[0] != "some":
                return None
            d = r[1]
            if best is None:
                best = d
            else:                                           # distance.min(min_distance): Rust f32::min returns the other operand for a NaN
                best = best if d != d else (d if best != best else (d if d < best else best))
    return best
