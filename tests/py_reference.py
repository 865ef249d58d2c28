"""A second, independent restatement of the reference's per-triangle path in
pure Python (floats are IEEE doubles, every operator rounds once -> the same
unfused arithmetic as Go/amd64).  Written from the Go source, not from the C
oracle, so that the two can check each other on small cases
(tests/test_oracle_cpu.py::test_c_oracle_matches_python_restatement).

Covers: Shader.Vertex (shader.go:70), the no-clip path of DrawTriangle
(context.go:370-389), drawClippedTriangle (context.go:316-349), rasterize
(context.go:151-281), InterpolateVertexes (vertex.go:18-47),
PhongShader.Fragment (shader.go:75-96), Color.NRGBA (color.go:56-63), the alpha
blend (context.go:256-267) and Go's math.Pow.  Triangle-index order.
"""
import math

MAXF = 1.7976931348623157e308


def go_int(x):
    if x != x or not (-9223372036854775808.0 < x < 9223372036854775808.0):
        return -(1 << 63)
    return int(x)


def go_pow(x, y):
    if y == 0 or x == 1:
        return 1.0
    if y == 1:
        return x
    if y == 0.5:
        return math.sqrt(x)
    yi, yf = float(int(abs(y))), abs(y) - int(abs(y))
    a1, ae = 1.0, 0
    if yf != 0:
        if yf > 0.5:
            yf -= 1
            yi += 1
        a1 = math.exp(yf * math.log(x))
    x1, xe = math.frexp(x)
    i = int(yi)
    while i != 0:
        if xe < -(1 << 12) or (1 << 12) < xe:
            ae += xe
            break
        if i & 1:
            a1 *= x1
            ae += xe
        x1 *= x1
        xe <<= 1
        if x1 < .5:
            x1 += x1
            xe -= 1
        i >>= 1
    if y < 0:
        a1 = 1 / a1
        ae = -ae
    return math.ldexp(a1, ae)


def mul_position_w(m, p):
    return (m[0] * p[0] + m[1] * p[1] + m[2] * p[2] + m[3],
            m[4] * p[0] + m[5] * p[1] + m[6] * p[2] + m[7],
            m[8] * p[0] + m[9] * p[1] + m[10] * p[2] + m[11],
            m[12] * p[0] + m[13] * p[1] + m[14] * p[2] + m[15])


def edge(a, b, c):
    return (b[0] - c[0]) * (a[1] - c[1]) - (b[1] - c[1]) * (a[0] - c[0])


def normalize(v):
    r = 1 / math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return (v[0] * r, v[1] * r, v[2] * r)


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def interp(a, b, c, B):
    return tuple((((0.0 + a[k] * B[0]) + b[k] * B[1]) + c[k] * B[2]) * B[3] for k in range(len(a)))


def clamp01(x):
    return 0.0 if x < 0 else (1.0 if x > 1 else x)


class PyContext:
    def __init__(self, w, h):
        self.W, self.H = w, h
        self.color = [[0, 0, 0, 0] for _ in range(w * h)]
        self.depth = [MAXF] * (w * h)
        self.read_depth = self.write_depth = self.write_color = self.alpha_blend = True
        self.cull_back = True
        self.depth_bias = 0.0

    def phong(self, sh, pos, nrm, col):
        light = list(sh["ambient"])
        color = col
        if tuple(sh["object"]) != (0, 0, 0, 0):
            color = tuple(sh["object"])
        ld = sh["light"]
        diffuse = max(dot(nrm, ld), 0.0)
        light = [light[k] + sh["diffuse"][k] * diffuse for k in range(4)]
        if diffuse > 0 and sh["specular_power"] > 0:
            cam = normalize(tuple(sh["camera"][k] - pos[k] for k in range(3)))
            i = (-ld[0], -ld[1], -ld[2])
            t = 2 * dot(nrm, i)
            refl = tuple(i[k] - nrm[k] * t for k in range(3))
            spec = max(dot(cam, refl), 0.0)
            if spec > 0:
                spec = go_pow(spec, sh["specular_power"])
                light = [light[k] + sh["specular"][k] * spec for k in range(4)]
        out = [min(color[k] * light[k], 1.0) for k in range(4)]
        out[3] = color[3]
        return out

    def draw_triangle(self, sh, P, N, C):
        """P,N: 3x3 lists; C: 3x4.  Returns (total, updated)."""
        W, H = self.W, self.H
        out = [mul_position_w(sh["matrix"], p) for p in P]
        for o in out:
            x, y, z, w = o
            if x < -w or x > w or y < -w or y > w or z < -w or z > w:
                raise NotImplementedError("clip path not restated here")
        ndc = [(o[0] / o[3], o[1] / o[3], o[2] / o[3]) for o in out]
        idx = [0, 1, 2]
        a = (ndc[1][0] - ndc[0][0]) * (ndc[2][1] - ndc[0][1]) - (ndc[2][0] - ndc[0][0]) * (ndc[1][1] - ndc[0][1])
        if a < 0:
            idx = [2, 1, 0]
        if self.cull_back and a <= 0:
            # a keeps its sign after the swap (context.go:324-336): CullBack + FaceCCW drops a <= 0
            return (0, 0)
        ndc = [ndc[i] for i in idx]
        w2, h2 = W / 2, H / 2
        s = [(w2 * n[0] + 0 * n[1] + 0 * n[2] + w2, 0 * n[0] + -h2 * n[1] + 0 * n[2] + h2,
              0 * n[0] + 0 * n[1] + 0.5 * n[2] + 0.5) for n in ndc]
        vP, vN, vC, vW = [P[i] for i in idx], [N[i] for i in idx], [C[i] for i in idx], [out[i][3] for i in idx]
        s0, s1, s2 = s
        x0 = go_int(math.floor(min(s0[0], min(s1[0], s2[0])))); x1 = go_int(math.ceil(max(s0[0], max(s1[0], s2[0]))))
        y0 = go_int(math.floor(min(s0[1], min(s1[1], s2[1])))); y1 = go_int(math.ceil(max(s0[1], max(s1[1], s2[1]))))
        p = (x0 + 0.5, y0 + 0.5)
        w00, w01, w02 = edge(s1, s2, p), edge(s2, s0, p), edge(s0, s1, p)
        a01, b01 = s1[1] - s0[1], s0[0] - s1[0]
        a12, b12 = s2[1] - s1[1], s1[0] - s2[0]
        a20, b20 = s0[1] - s2[1], s2[0] - s0[0]

        def recip(v):
            return math.copysign(math.inf, v) if v == 0 else 1 / v
        ra = recip(edge(s0, s1, s2))
        r0, r1, r2 = 1 / vW[0], 1 / vW[1], 1 / vW[2]
        ra12, ra20, ra01 = recip(a12), recip(a20), recip(a01)
        total = updated = 0
        for y in range(y0, y1 + 1):
            d = 0.0
            d0, d1, d2 = -w00 * ra12, -w01 * ra20, -w02 * ra01
            if w00 < 0 and d0 > d: d = d0
            if w01 < 0 and d1 > d: d = d1
            if w02 < 0 and d2 > d: d = d2
            d = float(go_int(d))
            if d < 0: d = 0.0
            w0, w1, w2_ = w00 + a12 * d, w01 + a20 * d, w02 + a01 * d
            was_inside = False
            x = x0 + go_int(d)
            while x <= x1:
                b0, b1, b2 = w0 * ra, w1 * ra, w2_ * ra
                w0 += a12; w1 += a20; w2_ += a01
                xx = x
                x += 1
                if b0 < 0 or b1 < 0 or b2 < 0:
                    if was_inside:
                        break
                    continue
                was_inside = True
                i = y * W + xx
                if i < 0 or i >= W * H:        # context.go:224: the index only -- x is never range-checked
                    continue
                total += 1
                z = b0 * s0[2] + b1 * s1[2] + b2 * s2[2]
                bz = z + self.depth_bias
                if self.read_depth and bz > self.depth[i]:
                    continue
                B = [b0 * r0, b1 * r1, b2 * r2, 0.0]
                B[3] = 1 / (B[0] + B[1] + B[2])
                pos = interp(vP[0], vP[1], vP[2], B)
                nrm = normalize(interp(vN[0], vN[1], vN[2], B))
                col = interp(vC[0], vC[1], vC[2], B)
                color = self.phong(sh, pos, nrm, col)
                if color == [0, 0, 0, 0]:
                    continue
                if bz <= self.depth[i] or not self.read_depth:
                    updated += 1
                    if self.write_depth:
                        self.depth[i] = z
                    if self.write_color:
                        c8 = [go_int(clamp01(c) * 255.0) & 0xff for c in color]
                        if self.alpha_blend and color[3] < 1:
                            sa = c8[3] * 0x101
                            src = [(c8[k] * 0x101) * c8[3] // 0xff for k in range(3)] + [sa]
                            aa = (0xffff - sa) * 0x101
                            self.color[i] = [((self.color[i][k] * aa // 0xffff + src[k]) >> 8) & 0xff for k in range(4)]
                        elif 0 <= xx < W:          # SetNRGBA's bounds check, context.go:269
                            self.color[i] = c8
            w00 += b12; w01 += b20; w02 += b01
        return (total, updated)
