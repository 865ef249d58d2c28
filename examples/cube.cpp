// examples/cube.cpp -- the reference's "hello world" flow written against the C++
// mirror (include/fauxgl.hpp): build a mesh, NewContext, clear, set a Phong
// shader, DrawMesh, then an outline pass with DrawLines, LineWidth and DepthBias,
// and read the image back.  Writes a raw dump that tests/test_cpp_host.py compares
// bit for bit with the CPU oracle.
//
//   g++ -std=c++17 -O2 -ffp-contract=off -Iinclude examples/cube.cpp \
//       -Lfauxgl_b200 -lfauxgl_b200 -Wl,-rpath,$PWD/fauxgl_b200 -o cube
#include <cstdio>
#include <cstdlib>

#include "fauxgl.hpp"

using namespace fauxgl;

int main(int argc, char **argv) {
    const char *out_path = argc > 1 ? argv[1] : "cube.raw";
    const int width = 320, height = 200;
    try {
        // shapes.go:17-39 NewCube, scaled to +-0.5
        const Vector v[8] = {{-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1}, {1, -1, -1}, {1, -1, 1}, {1, 1, -1}, {1, 1, 1}};
        const int idx[12][3] = {{3, 5, 7}, {5, 3, 1}, {0, 6, 4}, {6, 0, 2}, {0, 5, 1}, {5, 0, 4},
                                {5, 6, 7}, {6, 5, 4}, {6, 3, 7}, {3, 6, 2}, {0, 3, 2}, {3, 0, 1}};
        Mesh mesh;
        for (auto &t : idx) mesh.AddTriangle(v[t[0]].MulScalar(0.5), v[t[1]].MulScalar(0.5), v[t[2]].MulScalar(0.5));
        const int edges[12][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {0, 4}, {2, 6}, {1, 5}, {3, 7}};
        for (auto &e : edges) mesh.AddLine(v[e[0]].MulScalar(0.5), v[e[1]].MulScalar(0.5));

        auto context = NewContext(width, height);
        context->ClearColorBufferWith(HexColor("#24221F"));

        const Vector eye = V(2, 1.5, 1.2), center = V(0, 0, 0), up = V(0, 0, 1);
        const double aspect = double(width) / double(height);
        const Matrix matrix = LookAt(eye, center, up).Perspective(40, aspect, 1, 10);
        Shader shader = NewPhongShader(matrix, V(0.75, 0.5, 1).Normalize(), eye);
        shader.ObjectColor = HexColor("#468966");
        context->shader = shader;
        RasterizeInfo info1 = context->DrawTriangles(mesh);

        context->shader = NewSolidColorShader(matrix, Black);
        context->LineWidth = 3;
        context->DepthBias = -1e-4;
        RasterizeInfo info2 = context->DrawLines(mesh);

        std::vector<uint8_t> image = context->Image();
        std::vector<double> depth = context->DepthBuffer();
        FILE *f = std::fopen(out_path, "wb");
        if (!f) { std::perror(out_path); return 2; }
        const int32_t hdr[2] = {width, height};
        const uint64_t infos[4] = {info1.TotalPixels, info1.UpdatedPixels, info2.TotalPixels, info2.UpdatedPixels};
        const double light[3] = {shader.LightDirection.X, shader.LightDirection.Y, shader.LightDirection.Z};
        std::fwrite(hdr, sizeof hdr, 1, f);
        std::fwrite(matrix.m, sizeof matrix.m, 1, f);
        std::fwrite(light, sizeof light, 1, f);
        std::fwrite(infos, sizeof infos, 1, f);
        std::fwrite(image.data(), 1, image.size(), f);
        std::fwrite(depth.data(), sizeof(double), depth.size(), f);
        std::fclose(f);
        std::printf("triangles: %llu/%llu  lines: %llu/%llu\n", (unsigned long long)info1.TotalPixels,
                    (unsigned long long)info1.UpdatedPixels, (unsigned long long)info2.TotalPixels,
                    (unsigned long long)info2.UpdatedPixels);
    } catch (const Error &e) {
        std::fprintf(stderr, "%s (status %d)\n", e.what(), e.status);
        return 1;
    }
    return 0;
}
