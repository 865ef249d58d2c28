// fauxgl.hpp -- C++ host-side mirror of fogleman/fauxgl's render API over the
// C ABI of fauxgl_b200.h (header only).
//
// The reference is a Go package; the build image has no Go toolchain, so besides
// the cgo shim (go/fauxgl) the host side is provided in C++ with the
// reference's names and argument meaning: V, Vector, Matrix (LookAt, Perspective,
// ...), Color/HexColor, Mesh, NewPhongShader/NewSolidColorShader/NewTextureShader,
// NewContext, ClearColorBufferWith, ClearDepthBuffer, DrawMesh/DrawTriangles/
// DrawLines, Image().  Matrix/vector expressions keep the reference's evaluation
// order (compile host code with -ffp-contract=off to keep them unfused).
//
// Everything that renders goes through libfauxgl_b200.so (sm_100a CUDA kernels).
// There is no CPU fallback: errors surface as fauxgl::Error.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "fauxgl_b200.h"

namespace fauxgl {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &m) : std::runtime_error("fauxgl_b200: " + m), status(s) {}
};
inline void check(int rc, const fgl_ctx *ctx = nullptr) {
    if (rc != 0) throw Error(rc, fgl_last_error(ctx));
}

// ---- vector.go ---------------------------------------------------------------------
struct Vector {
    double X = 0, Y = 0, Z = 0;
    Vector Add(Vector b) const { return {X + b.X, Y + b.Y, Z + b.Z}; }
    Vector Sub(Vector b) const { return {X - b.X, Y - b.Y, Z - b.Z}; }
    Vector MulScalar(double b) const { return {X * b, Y * b, Z * b}; }
    double Dot(Vector b) const { return X * b.X + Y * b.Y + Z * b.Z; }
    Vector Cross(Vector b) const { return {Y * b.Z - Z * b.Y, Z * b.X - X * b.Z, X * b.Y - Y * b.X}; }
    Vector Normalize() const {
        double r = 1 / std::sqrt(X * X + Y * Y + Z * Z);
        return {X * r, Y * r, Z * r};
    }
    Vector Negate() const { return {-X, -Y, -Z}; }
};
inline Vector V(double x, double y, double z) { return {x, y, z}; }
inline double Radians(double degrees) { return degrees * M_PI / 180; }

// ---- color.go ------------------------------------------------------------------------
struct Color {
    double R = 0, G = 0, B = 0, A = 0;
    Color Alpha(double a) const { return {R, G, B, a}; }
    void NRGBA(uint8_t out[4]) const {  // color.go:56-63
        auto q = [](double x) { x = x < 0 ? 0 : (x > 1 ? 1 : x); return (uint8_t)(x * 255); };
        out[0] = q(R); out[1] = q(G); out[2] = q(B); out[3] = q(A);
    }
};
const Color Discard{0, 0, 0, 0}, Transparent{0, 0, 0, 0}, Black{0, 0, 0, 1}, White{1, 1, 1, 1};
inline Color Gray(double x) { return {x, x, x, 1}; }
inline Color HexColor(std::string x) {  // color.go:31-54 (6- and 8-digit forms)
    if (!x.empty() && x[0] == '#') x = x.substr(1);
    unsigned r = 0, g = 0, b = 0, a = 255;
    if (x.size() == 6) std::sscanf(x.c_str(), "%02x%02x%02x", &r, &g, &b);
    else if (x.size() == 8) std::sscanf(x.c_str(), "%02x%02x%02x%02x", &r, &g, &b, &a);
    return {r / 255.0, g / 255.0, b / 255.0, a / 255.0};
}

// ---- matrix.go ---------------------------------------------------------------------------
struct Matrix {
    double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // row-major X00..X33
    Matrix Mul(const Matrix &b) const {  // matrix.go:188-207
        Matrix r;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                r.m[4 * i + j] = m[4 * i] * b.m[j] + m[4 * i + 1] * b.m[4 + j] + m[4 * i + 2] * b.m[8 + j] + m[4 * i + 3] * b.m[12 + j];
        return r;
    }
    Matrix Perspective(double fovy, double aspect, double near, double far) const;  // left-multiplies, matrix.go:167
    Matrix Translate(Vector v) const;
    Matrix Scale(Vector v) const;
    Matrix Rotate(Vector v, double a) const;
};
inline Matrix Identity() { return {}; }
inline Matrix Translate(Vector v) { Matrix r; r.m[3] = v.X; r.m[7] = v.Y; r.m[11] = v.Z; return r; }
inline Matrix Scale(Vector v) { Matrix r; r.m[0] = v.X; r.m[5] = v.Y; r.m[10] = v.Z; return r; }
inline Matrix Rotate(Vector v, double a) {  // matrix.go:36-46
    v = v.Normalize();
    double s = std::sin(a), c = std::cos(a), k = 1 - c;
    Matrix r;
    double t[16] = {k * v.X * v.X + c, k * v.X * v.Y + v.Z * s, k * v.Z * v.X - v.Y * s, 0,
                    k * v.X * v.Y - v.Z * s, k * v.Y * v.Y + c, k * v.Y * v.Z + v.X * s, 0,
                    k * v.Z * v.X + v.Y * s, k * v.Y * v.Z - v.X * s, k * v.Z * v.Z + c, 0, 0, 0, 0, 1};
    std::memcpy(r.m, t, sizeof t);
    return r;
}
inline Matrix Frustum(double l, double r, double b, double t, double n, double f) {  // matrix.go:69-79
    double t1 = 2 * n, t2 = r - l, t3 = t - b, t4 = f - n;
    Matrix o;
    double v[16] = {t1 / t2, 0, (r + l) / t2, 0, 0, t1 / t3, (t + b) / t3, 0, 0, 0, (-f - n) / t4, (-t1 * f) / t4, 0, 0, -1, 0};
    std::memcpy(o.m, v, sizeof v);
    return o;
}
inline Matrix Perspective(double fovy, double aspect, double near, double far) {  // matrix.go:89-93
    double ymax = near * std::tan(fovy * M_PI / 360), xmax = ymax * aspect;
    return Frustum(-xmax, xmax, -ymax, ymax, near, far);
}
inline Matrix Orthographic(double l, double r, double b, double t, double n, double f) {  // matrix.go:81-87
    Matrix o;
    double v[16] = {2 / (r - l), 0, 0, -(r + l) / (r - l), 0, 2 / (t - b), 0, -(t + b) / (t - b),
                    0, 0, -2 / (f - n), -(f + n) / (f - n), 0, 0, 0, 1};
    std::memcpy(o.m, v, sizeof v);
    return o;
}
inline Matrix LookAt(Vector eye, Vector center, Vector up) {  // matrix.go:95-105
    Vector z = eye.Sub(center).Normalize(), x = up.Cross(z).Normalize(), y = z.Cross(x);
    Matrix o;
    double v[16] = {x.X, x.Y, x.Z, -x.Dot(eye), y.X, y.Y, y.Z, -y.Dot(eye), z.X, z.Y, z.Z, -z.Dot(eye), 0, 0, 0, 1};
    std::memcpy(o.m, v, sizeof v);
    return o;
}
inline Matrix Matrix::Perspective(double fovy, double aspect, double n, double f) const { return fauxgl::Perspective(fovy, aspect, n, f).Mul(*this); }
inline Matrix Matrix::Translate(Vector v) const { return fauxgl::Translate(v).Mul(*this); }
inline Matrix Matrix::Scale(Vector v) const { return fauxgl::Scale(v).Mul(*this); }
inline Matrix Matrix::Rotate(Vector v, double a) const { return fauxgl::Rotate(v, a).Mul(*this); }

// ---- mesh.go (attribute-major host mesh: [T][3][k] float64) -------------------------------------
struct Mesh {
    std::vector<double> position, normal, texture, color;      // triangles: 9,9,9,12 doubles each
    std::vector<double> lposition, lnormal, ltexture, lcolor;  // lines: 6,6,6,8 doubles each
    uint64_t generation = 0;
    size_t NumTriangles() const { return position.size() / 9; }
    size_t NumLines() const { return lposition.size() / 6; }
    void Invalidate() { generation++; }
    void AddTriangle(Vector a, Vector b, Vector c) {  // NewTriangleForPoints + FixNormals (triangle.go:13-18,46-58)
        Vector n = b.Sub(a).Cross(c.Sub(a)).Normalize();
        for (Vector p : {a, b, c}) {
            position.insert(position.end(), {p.X, p.Y, p.Z});
            normal.insert(normal.end(), {n.X, n.Y, n.Z});
            texture.insert(texture.end(), {0, 0, 0});
            color.insert(color.end(), {0, 0, 0, 0});
        }
        generation++;
    }
    void AddLine(Vector a, Vector b) {
        for (Vector p : {a, b}) {
            lposition.insert(lposition.end(), {p.X, p.Y, p.Z});
            lnormal.insert(lnormal.end(), {0, 0, 0});
            ltexture.insert(ltexture.end(), {0, 0, 0});
            lcolor.insert(lcolor.end(), {0, 0, 0, 0});
        }
        generation++;
    }
};

// ---- texture.go / shader.go ------------------------------------------------------------------------
struct ImageTexture {
    int Width = 0, Height = 0, Format = FGL_TEX_RGBA;
    std::vector<uint8_t> Pixels;  // RGBA8 rows
};
struct Shader {
    int kind = FGL_SHADER_SOLID;
    Matrix matrix;
    Vector LightDirection, CameraPosition;
    Color ObjectColor = Discard, AmbientColor{0.2, 0.2, 0.2, 1}, DiffuseColor{0.8, 0.8, 0.8, 1}, SpecularColor{1, 1, 1, 1};
    double SpecularPower = 32;
    Color SolidColor{1, 0, 1, 1};
    std::shared_ptr<ImageTexture> Texture;
};
inline Shader NewSolidColorShader(const Matrix &m, Color c) { Shader s; s.kind = FGL_SHADER_SOLID; s.matrix = m; s.SolidColor = c; return s; }
inline Shader NewTextureShader(const Matrix &m, std::shared_ptr<ImageTexture> t) { Shader s; s.kind = FGL_SHADER_TEXTURE; s.matrix = m; s.Texture = std::move(t); return s; }
inline Shader NewPhongShader(const Matrix &m, Vector light, Vector camera) {  // shader.go:61-68
    Shader s; s.kind = FGL_SHADER_PHONG; s.matrix = m; s.LightDirection = light; s.CameraPosition = camera; return s;
}

enum Face { FaceCW = FGL_FACE_CW, FaceCCW = FGL_FACE_CCW };
enum Cull { CullNone = FGL_CULL_NONE, CullFront = FGL_CULL_FRONT, CullBack = FGL_CULL_BACK };
struct RasterizeInfo {
    uint64_t TotalPixels = 0, UpdatedPixels = 0;
    RasterizeInfo Add(RasterizeInfo o) const { return {TotalPixels + o.TotalPixels, UpdatedPixels + o.UpdatedPixels}; }
};

// ---- context.go -------------------------------------------------------------------------------------
class Context {
  public:
    int Width, Height;
    Color ClearColor = Transparent;
    Shader shader = NewSolidColorShader(Identity(), Color{1, 0, 1, 1});
    bool ReadDepth = true, WriteDepth = true, WriteColor = true, AlphaBlend = true, Wireframe = false;
    Face FrontFace = FaceCCW;
    Cull cull = CullBack;
    double LineWidth = 2, DepthBias = 0;
    bool XGuard = false;  // not in the reference: true drops fragments with x outside [0, Width) (fgl_state.x_guard)

    Context(int width, int height, int device = 0) : Width(width), Height(height) { check(fgl_context_create(width, height, device, &h_)); }
    ~Context() {
        if (mesh_) fgl_mesh_destroy(mesh_);
        if (tex_) fgl_texture_destroy(tex_);
        fgl_context_destroy(h_);
    }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;

    void ClearColorBufferWith(Color c) { uint8_t p[4]; c.NRGBA(p); check(fgl_clear_color(h_, p), h_); }
    void ClearColorBuffer() { ClearColorBufferWith(ClearColor); }
    void ClearDepthBufferWith(double v) { check(fgl_clear_depth(h_, v), h_); }
    void ClearDepthBuffer() { ClearDepthBufferWith(1.7976931348623157e308); }

    RasterizeInfo DrawTriangles(const Mesh &mesh) { return draw(mesh, true, false); }
    RasterizeInfo DrawLines(const Mesh &mesh) { return draw(mesh, false, true); }
    RasterizeInfo DrawMesh(const Mesh &mesh) { return draw(mesh, true, true); }

    std::vector<uint8_t> Image() {  // ColorBuffer.Pix, NRGBA8
        std::vector<uint8_t> out((size_t)Width * Height * 4);
        check(fgl_read_color(h_, out.data(), 0), h_);
        return out;
    }
    std::vector<double> DepthBuffer() {
        std::vector<double> out((size_t)Width * Height);
        check(fgl_read_depth(h_, out.data()), h_);
        return out;
    }
    std::vector<uint16_t> DepthImage() {  // context.go:87-117, Gray16
        std::vector<uint16_t> out((size_t)Width * Height);
        check(fgl_depth_image(h_, out.data()), h_);
        return out;
    }
    // One RasterizeInfo per line / triangle, as a loop of Context.DrawLine / DrawTriangle would return
    // (context.go:351-389; examples/silhouette.go:163-166), in one launch sequence.
    std::vector<RasterizeInfo> DrawLinesEach(const Mesh &mesh) { return draw_each(mesh, true); }
    std::vector<RasterizeInfo> DrawTrianglesEach(const Mesh &mesh) { return draw_each(mesh, false); }
    std::vector<uint8_t> Resolve(int factor) {  // resize.Resize(w/f, h/f, Image(), resize.Bilinear)
        std::vector<uint8_t> out((size_t)(Width / factor) * (Height / factor) * 4);
        check(fgl_resolve(h_, factor, out.data()), h_);
        return out;
    }
    fgl_ctx *handle() { return h_; }

  private:
    void setup(const Mesh &mesh, fgl_state &st, fgl_shader &sh) {
        if (mesh_src_ != &mesh || mesh_gen_ != mesh.generation) upload(mesh);
        st.read_depth = ReadDepth; st.write_depth = WriteDepth; st.write_color = WriteColor;
        st.alpha_blend = AlphaBlend; st.wireframe = Wireframe; st.front_face = FrontFace; st.cull = cull;
        st.line_width = LineWidth; st.depth_bias = DepthBias; st.x_guard = XGuard;
        sh.kind = shader.kind;
        std::memcpy(sh.matrix, shader.matrix.m, sizeof sh.matrix);
        auto p3 = [](double *d, Vector v) { d[0] = v.X; d[1] = v.Y; d[2] = v.Z; };
        auto p4 = [](double *d, Color c) { d[0] = c.R; d[1] = c.G; d[2] = c.B; d[3] = c.A; };
        p3(sh.light, shader.LightDirection); p3(sh.camera, shader.CameraPosition);
        p4(sh.object, shader.ObjectColor); p4(sh.ambient, shader.AmbientColor); p4(sh.diffuse, shader.DiffuseColor);
        p4(sh.specular, shader.SpecularColor); p4(sh.color, shader.SolidColor);
        sh.specular_power = shader.SpecularPower;
        if (shader.Texture) {
            if (tex_src_ != shader.Texture.get()) {
                if (tex_) fgl_texture_destroy(tex_);
                tex_ = nullptr;
                const ImageTexture &t = *shader.Texture;
                check(fgl_texture_create(h_, t.Pixels.data(), t.Width, t.Height, t.Format, &tex_), h_);
                tex_src_ = shader.Texture.get();
            }
            sh.texture = tex_;
        }
    }
    RasterizeInfo draw(const Mesh &mesh, bool tris, bool lines) {
        fgl_state st{};
        fgl_shader sh{};
        setup(mesh, st, sh);
        RasterizeInfo result;
        fgl_raster_info info{};
        if (tris && mesh.NumTriangles()) {
            check(fgl_draw_triangles(h_, &st, &sh, mesh_, 0, mesh.NumTriangles(), &info), h_);
            result = result.Add({info.total_pixels, info.updated_pixels});
        }
        if (lines && mesh.NumLines()) {
            check(fgl_draw_lines(h_, &st, &sh, mesh_, 0, mesh.NumLines(), &info), h_);
            result = result.Add({info.total_pixels, info.updated_pixels});
        }
        return result;
    }
    std::vector<RasterizeInfo> draw_each(const Mesh &mesh, bool lines) {
        fgl_state st{};
        fgl_shader sh{};
        setup(mesh, st, sh);
        const size_t n = lines ? mesh.NumLines() : mesh.NumTriangles();
        std::vector<fgl_raster_info> infos(n);
        if (n) check((lines ? fgl_draw_lines_each : fgl_draw_triangles_each)(h_, &st, &sh, mesh_, 0, n, infos.data(), nullptr), h_);
        std::vector<RasterizeInfo> out(n);
        for (size_t i = 0; i < n; i++) out[i] = {infos[i].total_pixels, infos[i].updated_pixels};
        return out;
    }
    void upload(const Mesh &mesh) {
        fgl_mesh_desc d{};
        d.ntriangles = mesh.NumTriangles(); d.nlines = mesh.NumLines();
        d.position = mesh.position.data(); d.normal = mesh.normal.data(); d.texture = mesh.texture.data(); d.color = mesh.color.data();
        d.lposition = mesh.lposition.data(); d.lnormal = mesh.lnormal.data(); d.ltexture = mesh.ltexture.data(); d.lcolor = mesh.lcolor.data();
        uint64_t nt = 0, nl = 0;
        if (mesh_) fgl_mesh_counts(mesh_, &nt, &nl);
        if (mesh_ && mesh_src_ == &mesh && nt == d.ntriangles && nl == d.nlines) {
            check(fgl_mesh_update(h_, mesh_, &d), h_);
        } else {
            if (mesh_) fgl_mesh_destroy(mesh_);
            mesh_ = nullptr;
            check(fgl_mesh_create(h_, &d, &mesh_), h_);
        }
        mesh_src_ = &mesh;
        mesh_gen_ = mesh.generation;
    }
    fgl_ctx *h_ = nullptr;
    fgl_mesh *mesh_ = nullptr;
    const Mesh *mesh_src_ = nullptr;
    uint64_t mesh_gen_ = 0;
    fgl_tex *tex_ = nullptr;
    const ImageTexture *tex_src_ = nullptr;
};
inline std::unique_ptr<Context> NewContext(int width, int height) { return std::make_unique<Context>(width, height); }

}  // namespace fauxgl
