// cgo binding of libfauxgl_b200.so (include/fauxgl_b200.h) for the fauxgl
// package.  One call per DrawTriangles/DrawLines, never per primitive.
//
// cgo may not pass memory that contains Go pointers, and []*Triangle is a slice
// of pointers: the triangles are gathered into flat float64 arrays
// ([T][3][k], the layout fgl_mesh_desc takes) before the call.  That gather is
// the only per-triangle work left on the host and happens once per mesh upload,
// not per frame.
//
// Build: CGO_CFLAGS=-I<repo>/include CGO_LDFLAGS="-L<repo>/fauxgl_b200 -lfauxgl_b200"
// (or the #cgo lines below with the library installed system-wide).
package fauxgl

/*
#cgo LDFLAGS: -lfauxgl_b200
#include <stdlib.h>
#include "fauxgl_b200.h"
*/
import "C"

import (
	"errors"
	"image"
	"math"
	"os"
	"strconv"
	"unsafe"
)

const (
	shaderSolid   = C.FGL_SHADER_SOLID
	shaderTexture = C.FGL_SHADER_TEXTURE
	shaderPhong   = C.FGL_SHADER_PHONG
)

type stateDesc struct {
	readDepth, writeDepth, writeColor, alphaBlend, wireframe bool
	frontFace, cull                                          int
	lineWidth, depthBias                                     float64
}

type shaderDesc struct {
	kind                               int
	matrix                             Matrix
	light, camera                      Vector
	object, ambient, diffuse, specular Color
	specularPower                      float64
	color                              Color
	texture                            *deviceTexture
}

type deviceContext struct{ h *C.fgl_ctx }
type deviceTexture struct{ h *C.fgl_tex }
type deviceMesh struct {
	h      *C.fgl_mesh
	nt, nl int
	fp     meshFingerprint
	stale  bool
}

func deviceFromEnv() int {
	if v, err := strconv.Atoi(os.Getenv("FAUXGL_DEVICE")); err == nil {
		return v
	}
	return 0
}

func lastError(ctx *C.fgl_ctx, rc C.int) error {
	if rc == 0 {
		return nil
	}
	return errors.New("fauxgl_b200: " + C.GoString(C.fgl_last_error(ctx)))
}

func newDeviceContext(w, h, device int) (*deviceContext, error) {
	var ctx *C.fgl_ctx
	if rc := C.fgl_context_create(C.int(w), C.int(h), C.int(device), &ctx); rc != 0 {
		return nil, lastError(nil, rc)
	}
	return &deviceContext{ctx}, nil
}

func (d *deviceContext) destroy() { C.fgl_context_destroy(d.h) }

func (d *deviceContext) clearColor(r, g, b, a uint8) error {
	c := [4]C.uint8_t{C.uint8_t(r), C.uint8_t(g), C.uint8_t(b), C.uint8_t(a)}
	return lastError(d.h, C.fgl_clear_color(d.h, &c[0]))
}

func (d *deviceContext) clearDepth(v float64) error {
	return lastError(d.h, C.fgl_clear_depth(d.h, C.double(v)))
}

func (d *deviceContext) readColor(pix []uint8, stride int) error {
	return lastError(d.h, C.fgl_read_color(d.h, (*C.uint8_t)(unsafe.Pointer(&pix[0])), C.size_t(stride)))
}

func (d *deviceContext) readDepth(depth []float64) error {
	return lastError(d.h, C.fgl_read_depth(d.h, (*C.double)(unsafe.Pointer(&depth[0]))))
}

func (d *deviceContext) writeColor(pix []uint8, stride int) error {
	return lastError(d.h, C.fgl_write_color(d.h, (*C.uint8_t)(unsafe.Pointer(&pix[0])), C.size_t(stride)))
}

func (d *deviceContext) writeDepth(depth []float64) error {
	return lastError(d.h, C.fgl_write_depth(d.h, (*C.double)(unsafe.Pointer(&depth[0]))))
}

func (d *deviceContext) depthImage(dst []uint16) error {
	return lastError(d.h, C.fgl_depth_image(d.h, (*C.uint16_t)(unsafe.Pointer(&dst[0]))))
}

func (d *deviceContext) resolve(factor int, dst []uint8) error {
	return lastError(d.h, C.fgl_resolve(d.h, C.int(factor), (*C.uint8_t)(unsafe.Pointer(&dst[0]))))
}

// newTexture uploads an ImageTexture.  *image.RGBA (what Go's PNG decoder yields
// for 8-bit RGB) and *image.NRGBA (8-bit RGBA) are passed through; MakeColor's
// RGBA() conversion for each is reproduced on the device (color.go:25-29).
func (d *deviceContext) newTexture(t *ImageTexture) (*deviceTexture, error) {
	var pix []uint8
	var stride, format int
	switch im := t.Image.(type) {
	case *image.RGBA:
		pix, stride, format = im.Pix, im.Stride, C.FGL_TEX_RGBA
	case *image.NRGBA:
		pix, stride, format = im.Pix, im.Stride, C.FGL_TEX_NRGBA
	default:
		return nil, errors.New("fauxgl: texture image type has no device path (need *image.RGBA or *image.NRGBA)")
	}
	if stride != 4*t.Width {
		packed := make([]uint8, 4*t.Width*t.Height)
		for y := 0; y < t.Height; y++ {
			copy(packed[4*t.Width*y:4*t.Width*(y+1)], pix[stride*y:])
		}
		pix = packed
	}
	var h *C.fgl_tex
	rc := C.fgl_texture_create(d.h, (*C.uint8_t)(unsafe.Pointer(&pix[0])), C.int(t.Width), C.int(t.Height), C.int(format), &h)
	if rc != 0 {
		return nil, lastError(d.h, rc)
	}
	return &deviceTexture{h}, nil
}

func (t *deviceTexture) destroy() { C.fgl_texture_destroy(t.h) }

// flatten gathers []*Triangle / []*Line into the per-attribute arrays of fgl_mesh_desc.
type flatMesh struct {
	pos, nrm, tex, col     []float64
	lpos, lnrm, ltex, lcol []float64
}

func putVertex(v *Vertex, i int, pos, nrm, tex, col []float64) {
	pos[3*i], pos[3*i+1], pos[3*i+2] = v.Position.X, v.Position.Y, v.Position.Z
	nrm[3*i], nrm[3*i+1], nrm[3*i+2] = v.Normal.X, v.Normal.Y, v.Normal.Z
	tex[3*i], tex[3*i+1], tex[3*i+2] = v.Texture.X, v.Texture.Y, v.Texture.Z
	col[4*i], col[4*i+1], col[4*i+2], col[4*i+3] = v.Color.R, v.Color.G, v.Color.B, v.Color.A
}

func flatten(mesh *Mesh) *flatMesh {
	nt, nl := len(mesh.Triangles), len(mesh.Lines)
	f := &flatMesh{
		pos: make([]float64, 9*nt), nrm: make([]float64, 9*nt), tex: make([]float64, 9*nt), col: make([]float64, 12*nt),
		lpos: make([]float64, 6*nl), lnrm: make([]float64, 6*nl), ltex: make([]float64, 6*nl), lcol: make([]float64, 8*nl),
	}
	for i, t := range mesh.Triangles {
		putVertex(&t.V1, 3*i, f.pos, f.nrm, f.tex, f.col)
		putVertex(&t.V2, 3*i+1, f.pos, f.nrm, f.tex, f.col)
		putVertex(&t.V3, 3*i+2, f.pos, f.nrm, f.tex, f.col)
	}
	for i, l := range mesh.Lines {
		putVertex(&l.V1, 2*i, f.lpos, f.lnrm, f.ltex, f.lcol)
		putVertex(&l.V2, 2*i+1, f.lpos, f.lnrm, f.ltex, f.lcol)
	}
	return f
}

func ptr(s []float64) *C.double {
	if len(s) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&s[0]))
}

func (f *flatMesh) desc(nt, nl int) C.fgl_mesh_desc {
	var d C.fgl_mesh_desc
	d.ntriangles, d.nlines = C.uint64_t(nt), C.uint64_t(nl)
	d.position, d.normal, d.texture, d.color = ptr(f.pos), ptr(f.nrm), ptr(f.tex), ptr(f.col)
	d.lposition, d.lnormal, d.ltexture, d.lcolor = ptr(f.lpos), ptr(f.lnrm), ptr(f.ltex), ptr(f.lcol)
	return d
}

func (d *deviceContext) newMesh(mesh *Mesh) (*deviceMesh, error) {
	f := flatten(mesh)
	desc := f.desc(len(mesh.Triangles), len(mesh.Lines))
	var h *C.fgl_mesh
	if rc := C.fgl_mesh_create(d.h, &desc, &h); rc != 0 {
		return nil, lastError(d.h, rc)
	}
	return &deviceMesh{h: h, nt: len(mesh.Triangles), nl: len(mesh.Lines)}, nil
}

func (m *deviceMesh) sameShape(mesh *Mesh) bool {
	return m.nt == len(mesh.Triangles) && m.nl == len(mesh.Lines)
}

func (m *deviceMesh) update(d *deviceContext, mesh *Mesh) error {
	f := flatten(mesh)
	desc := f.desc(m.nt, m.nl)
	return lastError(d.h, C.fgl_mesh_update(d.h, m.h, &desc))
}

// transform applies Mesh.Transform (mesh.go:167-175) to the device copy: an animate.go-style loop can rotate
// the resident mesh every frame without flattening and uploading it again.
func (m *deviceMesh) transform(d *deviceContext, matrix Matrix) error {
	mm := [16]C.double{
		C.double(matrix.X00), C.double(matrix.X01), C.double(matrix.X02), C.double(matrix.X03),
		C.double(matrix.X10), C.double(matrix.X11), C.double(matrix.X12), C.double(matrix.X13),
		C.double(matrix.X20), C.double(matrix.X21), C.double(matrix.X22), C.double(matrix.X23),
		C.double(matrix.X30), C.double(matrix.X31), C.double(matrix.X32), C.double(matrix.X33)}
	return lastError(d.h, C.fgl_mesh_transform(d.h, m.h, &mm[0]))
}

// smoothNormals is Mesh.SmoothNormals (mesh.go:105-120) on the device copy, bit-identical to the host loop.
func (m *deviceMesh) smoothNormals(d *deviceContext) error {
	return lastError(d.h, C.fgl_mesh_smooth_normals(d.h, m.h))
}

// smoothNormalsThreshold is Mesh.SmoothNormalsThreshold (mesh.go:90-103); math.Cos is taken here, in Go.
func (m *deviceMesh) smoothNormalsThreshold(d *deviceContext, radians float64) error {
	return lastError(d.h, C.fgl_mesh_smooth_normals_threshold(d.h, m.h, C.double(math.Cos(radians))))
}

func (m *deviceMesh) destroy() {
	if m.h != nil {
		C.fgl_mesh_destroy(m.h)
		m.h = nil
	}
}

// meshFingerprint is a cheap change detector for meshes mutated through Mesh
// methods between draws (mesh.Transform in examples/animate.go:66).
type meshFingerprint struct {
	nt, nl      int
	first, last Vertex
}

func fingerprint(mesh *Mesh) meshFingerprint {
	fp := meshFingerprint{nt: len(mesh.Triangles), nl: len(mesh.Lines)}
	if fp.nt > 0 {
		fp.first, fp.last = mesh.Triangles[0].V1, mesh.Triangles[fp.nt-1].V3
	} else if fp.nl > 0 {
		fp.first, fp.last = mesh.Lines[0].V1, mesh.Lines[fp.nl-1].V2
	}
	return fp
}

func cbool(b bool) C.int32_t {
	if b {
		return 1
	}
	return 0
}

func (s stateDesc) c() C.fgl_state {
	var st C.fgl_state
	st.read_depth, st.write_depth, st.write_color = cbool(s.readDepth), cbool(s.writeDepth), cbool(s.writeColor)
	st.alpha_blend, st.wireframe = cbool(s.alphaBlend), cbool(s.wireframe)
	st.front_face, st.cull = C.int32_t(s.frontFace), C.int32_t(s.cull)
	st.line_width, st.depth_bias = C.double(s.lineWidth), C.double(s.depthBias)
	return st
}

func (s shaderDesc) c() C.fgl_shader {
	var sh C.fgl_shader
	sh.kind = C.int32_t(s.kind)
	m := s.matrix
	for i, v := range [16]float64{m.X00, m.X01, m.X02, m.X03, m.X10, m.X11, m.X12, m.X13,
		m.X20, m.X21, m.X22, m.X23, m.X30, m.X31, m.X32, m.X33} {
		sh.matrix[i] = C.double(v)
	}
	put3 := func(dst *[3]C.double, v Vector) { dst[0], dst[1], dst[2] = C.double(v.X), C.double(v.Y), C.double(v.Z) }
	put4 := func(dst *[4]C.double, c Color) {
		dst[0], dst[1], dst[2], dst[3] = C.double(c.R), C.double(c.G), C.double(c.B), C.double(c.A)
	}
	put3(&sh.light, s.light)
	put3(&sh.camera, s.camera)
	put4(&sh.object, s.object)
	put4(&sh.ambient, s.ambient)
	put4(&sh.diffuse, s.diffuse)
	put4(&sh.specular, s.specular)
	put4(&sh.color, s.color)
	sh.specular_power = C.double(s.specularPower)
	if s.texture != nil {
		sh.texture = s.texture.h
	}
	return sh
}

func (d *deviceContext) drawTriangles(st stateDesc, sh shaderDesc, m *deviceMesh, first, count uint64) (RasterizeInfo, error) {
	cst, csh := st.c(), sh.c()
	var info C.fgl_raster_info
	rc := C.fgl_draw_triangles(d.h, &cst, &csh, m.h, C.uint64_t(first), C.uint64_t(count), &info)
	return RasterizeInfo{uint64(info.total_pixels), uint64(info.updated_pixels)}, lastError(d.h, rc)
}

func (d *deviceContext) drawLines(st stateDesc, sh shaderDesc, m *deviceMesh, first, count uint64) (RasterizeInfo, error) {
	cst, csh := st.c(), sh.c()
	var info C.fgl_raster_info
	rc := C.fgl_draw_lines(d.h, &cst, &csh, m.h, C.uint64_t(first), C.uint64_t(count), &info)
	return RasterizeInfo{uint64(info.total_pixels), uint64(info.updated_pixels)}, lastError(d.h, rc)
}

// drawEach fills out[i] with the RasterizeInfo of primitive i (RasterizeInfo is two
// uint64, the layout of fgl_raster_info).
func (d *deviceContext) drawEach(st stateDesc, sh shaderDesc, m *deviceMesh, lines bool, out []RasterizeInfo) error {
	cst, csh := st.c(), sh.c()
	infos := (*C.fgl_raster_info)(unsafe.Pointer(&out[0]))
	var rc C.int
	if lines {
		rc = C.fgl_draw_lines_each(d.h, &cst, &csh, m.h, 0, C.uint64_t(len(out)), infos, nil)
	} else {
		rc = C.fgl_draw_triangles_each(d.h, &cst, &csh, m.h, 0, C.uint64_t(len(out)), infos, nil)
	}
	return lastError(d.h, rc)
}
