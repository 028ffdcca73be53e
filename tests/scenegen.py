"""Seeded random scene graphs over the whole GlomeTrace vocabulary, for differential tests (oracle vs the scene-graph
machine on the CPU, oracle vs the CUDA kernels on a B200).  Every node kind of SURVEY.md section 8(a) appears:
Sphere, Triangle(Norm), Box, Plane (inside Intersections only: a bare plane has no finite bound, Bih.hs:319-322), Disc,
Cylinder, Cone, group, Void, Instance, Bih, Mesh, Difference, Intersection, Tex, Tag, NoShadow, OnlyShadow, Bound,
InnerBound; and every Material: Surface, Reflect, Refract, Warp, AdditiveLayers, Blend (+ stripe / perlin textures)."""
import numpy as np

import glome_b200 as G


class Gen:
    def __init__(self, seed, n_lights=2, with_mesh=True, with_warp=True):
        self.rng = np.random.default_rng(seed)
        self.b = G.SceneBuilder()
        b = self.b
        r = self.rng
        self.mats = []
        for _ in range(4):
            self.mats.append(b.mat_surface(r.uniform(0.1, 1, 3), 1.0, 0.2, r.uniform(0.4, 1), float(r.choice([0.0, 0.4])), 10.0))
        self.mats.append(b.mat_surface(r.uniform(0.1, 1, 3), 0.6, 0.2, 0.8, 0.0, 0.0))  # translucent: the texture fold goes on
        self.m_mirror = b.mat_reflect(0.8)
        self.m_glass = b.mat_refract(0.35, 0.8, 1.5)
        self.m_blend = b.mat_blend(self.mats[0], self.m_mirror, 0.3)
        self.m_add = b.mat_additive([self.mats[1], self.mats[4], self.m_mirror])
        self.m_add0 = b.mat_additive([])
        self.texs = [b.tex_uniform(m) for m in self.mats]
        self.texs += [b.tex_uniform(self.m_mirror), b.tex_uniform(self.m_glass), b.tex_uniform(self.m_blend),
                      b.tex_uniform(self.m_add), b.tex_uniform(self.m_add0),
                      b.tex_stripe_blend(self.mats[0], self.mats[2], (4, 8, 5)),
                      b.tex_perlin_blend(self.m_mirror, self.mats[3], 3.0)]
        self.n_lights = n_lights
        self.with_mesh = with_mesh
        self.with_warp = with_warp
        self.warp_mat = None

    # ---- leaves ----
    def pt(self, ext=4.0):
        return self.rng.uniform(-ext, ext, 3)

    def prim(self):
        r, b = self.rng, self.b
        k = int(r.integers(0, 7))
        c = self.pt()
        if k == 0:
            return b.sphere(c, r.uniform(0.3, 1.2))
        if k == 1:
            return b.box(c - r.uniform(0.2, 1.0, 3), c + r.uniform(0.2, 1.0, 3))
        if k == 2:
            return b.cone(c, r.uniform(0.3, 0.9), c + r.uniform(-1.5, 1.5, 3) + np.array([0, 1.0, 0]), r.uniform(0.0, 0.3))
        if k == 3:
            return b.cylinder(c, c + r.uniform(-1.5, 1.5, 3) + np.array([0.5, 0.5, 0]), r.uniform(0.2, 0.7))
        if k == 4:
            return b.triangle(c, c + r.uniform(-1.5, 1.5, 3), c + r.uniform(-1.5, 1.5, 3))
        if k == 5:
            n = [x / np.linalg.norm(x) for x in r.normal(size=(3, 3))]
            return b.trianglenorm(c, c + r.uniform(-1.5, 1.5, 3), c + r.uniform(-1.5, 1.5, 3), *n)
        n = r.normal(size=3)
        return b.disc(c, n / np.linalg.norm(n), r.uniform(0.4, 1.2))

    def solid_prim(self):  # something with an inside (CSG operands)
        r, b = self.rng, self.b
        k = int(r.integers(0, 4))
        c = self.pt(2.5)
        if k == 0:
            return b.sphere(c, r.uniform(0.6, 1.6))
        if k == 1:
            return b.box(c - r.uniform(0.4, 1.4, 3), c + r.uniform(0.4, 1.4, 3))
        if k == 2:
            return b.cylinder(c - np.array([0, 1.2, 0]), c + np.array([0.1, 1.2, 0.2]), r.uniform(0.4, 1.0))
        return b.cone(c - np.array([0, 1.0, 0]), r.uniform(0.5, 1.1), c + np.array([0, 1.0, 0]), r.uniform(0.05, 0.3))

    def wrap(self, item, p=0.6):
        r, b = self.rng, self.b
        while r.random() < p:
            k = int(r.integers(0, 10))
            if k < 5:
                item = b.tex(item, int(r.choice(self.texs)))
            elif k < 8:
                item = b.tag(item, int(r.integers(-5, 100000)))
            elif k == 8:
                item = b.noshadow(item)
            else:
                item = b.onlyshadow(item)
            p *= 0.5
        return item

    def xform(self, item):
        r = self.rng
        ax = r.normal(size=3)
        xs = [G.scale(r.uniform(0.6, 1.5, 3)), G.rotate(ax / np.linalg.norm(ax), float(r.uniform(-1, 1))), G.translate(r.uniform(-2, 2, 3))]
        return self.b.transform(item, xs[: int(r.integers(1, 4))])

    def mesh(self):
        r, b = self.rng, self.b
        g = int(r.integers(3, 7))
        xs, zs = np.meshgrid(np.linspace(-3, 3, g + 1), np.linspace(-3, 3, g + 1))
        ys = 0.4 * np.sin(xs * 1.3) * np.cos(zs * 0.9) - 2.0
        verts = np.stack([xs.ravel(), ys.ravel(), zs.ravel()], 1)
        norms = np.tile(np.array([[0.0, 1.0, 0.0]]), (len(verts), 1)) + 0.2 * r.normal(size=verts.shape)
        norms /= np.linalg.norm(norms, axis=1)[:, None]
        tris = []
        for j in range(g):
            for i in range(g):
                v00 = j * (g + 1) + i
                v10, v01, v11 = v00 + 1, v00 + g + 1, v00 + g + 2
                smooth = (i + j) % 3 != 0
                tx = int(r.integers(-1, 3))
                tg = int(r.integers(-1, 4))
                tris.append([v00, v01, v10] + ([v00, v01, v10] if smooth else [-1, -1, -1]) + [tx, tg])
                tris.append([v10, v01, v11] + ([v10, v01, v11] if smooth else [-1, -1, -1]) + [tx, tg])
        return b.mesh(verts, norms, np.array(tris, dtype=np.int32), [int(x) for x in r.choice(self.texs, 3)], [7, -3, 123456, 9])

    # ---- composites ----
    def node(self, depth):
        r, b = self.rng, self.b
        if depth <= 0:
            return self.wrap(self.xform(self.prim()) if r.random() < 0.3 else self.prim())
        k = int(r.integers(0, 12))
        if k <= 1:
            return self.wrap(self.prim())
        if k == 2:
            return self.wrap(b.group([self.node(depth - 1) for _ in range(int(r.integers(0, 5)))]))
        if k == 3:
            items = [self.node(depth - 1) for _ in range(int(r.integers(1, 9)))]
            return self.wrap(b.bih(items))
        if k == 4:
            return self.wrap(self.xform(self.node(depth - 1)))
        if k == 5:
            return self.wrap(b.difference(self.csg_operand(depth - 1), self.csg_operand(depth - 1)))
        if k == 6:
            n = int(r.integers(1, 5))
            parts = [self.csg_operand(depth - 1) for _ in range(n)]
            if r.random() < 0.5:
                nn = r.normal(size=3)
                parts.append(b.plane(self.pt(1.0), nn / np.linalg.norm(nn)))
            return self.wrap(b.intersection(parts))
        if k == 7:
            inner = self.node(depth - 1)
            bb = b.bound(inner)
            c, rad = 0.5 * (bb[:3] + bb[3:]), 0.5 * np.linalg.norm(bb[3:] - bb[:3]) * float(r.uniform(0.7, 1.1))
            return self.wrap(b.bound_object(b.sphere(c, max(rad, 0.1)), inner))
        if k == 8:
            return self.wrap(b.innerbound(self.solid_prim(), self.node(depth - 1)))
        if k == 9 and self.with_mesh:
            return self.wrap(self.mesh())
        if k == 10:
            return b.void()
        return self.wrap(b.bih([self.wrap(self.prim()) for _ in range(int(r.integers(2, 30)))]))

    def csg_operand(self, depth):
        r, b = self.rng, self.b
        k = int(r.integers(0, 8))
        if depth <= 0 or k <= 3:
            s = self.solid_prim()
            return self.wrap(self.xform(s) if r.random() < 0.3 else s, 0.4)
        if k == 4:
            return self.wrap(b.difference(self.csg_operand(depth - 1), self.csg_operand(depth - 1)), 0.4)
        if k == 5:
            return self.wrap(b.intersection([self.csg_operand(depth - 1) for _ in range(int(r.integers(1, 4)))]), 0.4)
        if k == 6:
            return self.wrap(b.group([self.csg_operand(depth - 1) for _ in range(int(r.integers(1, 4)))]), 0.4)
        return self.wrap(b.bih([self.wrap(self.solid_prim(), 0.3) for _ in range(int(r.integers(1, 8)))]), 0.4)

    def scene(self, depth=3, top=6):
        r, b = self.rng, self.b
        for _ in range(self.n_lights):
            b.light(r.uniform(-12, 12, 3) + np.array([0, 14, 0]), r.uniform(40, 160, 3))
        items = [self.node(depth) for _ in range(top)]
        # a floor so that most rays end on something
        items.append(b.tex(b.box((-9, -4.5, -9), (9, -4, 9)), int(self.rng.choice(self.texs))))
        if self.with_warp:
            frame = b.tag(b.tex(b.box((-1, -1, -0.1), (1, 1, 0.1)), self.texs[0]), 77)
            self.warp_mat = b.mat_warp(frame, -1, 0, G.compose([G.rotate((1, 0, 0), 0.4), G.translate((1.0, 3.0, -2.0))]))
            items.append(b.transform(b.group([b.tex(b.box((-0.8, -0.8, -0.05), (0.8, 0.8, 0.05)), b.tex_uniform(self.warp_mat))]),
                                     [G.translate((3.5, 0.5, 2.0))]))
        root = b.bih(items) if r.random() < 0.7 else b.group(items)
        if self.warp_mat is not None:
            b.mat_warp_set_scene(self.warp_mat, root)
        cam = G.camera(r.uniform(-2, 2, 3) + np.array([0, 3.0, 13.0]), (0, 0, 0), (0, 1, 0), 50)
        return root, cam


def random_scene(seed, **kw):
    """-> (builder, flat scene, camera)"""
    depth = kw.pop("depth", 3)
    top = kw.pop("top", 6)
    g = Gen(seed, **kw)
    root, cam = g.scene(depth, top)
    return g.b, g.b.flatten(root), cam


def query_rays(cam, w, h, seed, nrand=1500, ext=6.0):
    ys, xs = np.mgrid[0:h, 0:w]
    rays = G.camera_rays(cam, w, h, xs.ravel(), ys.ravel())
    rng = np.random.default_rng(seed)
    o = rng.uniform(-ext, ext, size=(nrand, 3))
    d = rng.normal(size=(nrand, 3))
    d /= np.sqrt((d * d).sum(1))[:, None]
    return np.vstack([rays, np.hstack([o, d])])
