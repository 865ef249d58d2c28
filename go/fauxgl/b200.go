// cgo binding of libfauxgl_b200.so (include/fauxgl_b200.h) for the fauxgl
// package.  One call per DrawTriangles/DrawLines, never per primitive.
//
// cgo pointer rules.  C may not be handed Go memory that itself contains Go
// pointers -- neither []*Triangle nor a Go-allocated fgl_mesh_desc whose fields
// point into Go slices (cgocheck panics with "cgo argument has Go pointer to
// unpinned Go pointer").  So the vertices are gathered into ONE block of C
// memory per device mesh (page-locked, fgl_host_alloc: no Go pointers inside,
// and H2D copies from it run at full PCIe rate), and the fgl_mesh_desc passed
// to the library holds only C pointers into that block.  Plain []uint8 /
// []float64 buffers (Pix, DepthBuffer) contain no pointers and are passed
// directly, pinned by cgo for the duration of the call.
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
	xGuard                                                   bool
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
	h       *C.fgl_mesh
	nt, nl  int
	host    unsafe.Pointer // C memory (fgl_host_alloc): the flattened attributes, [pos|nrm|tex|col] triangles then lines
	hostLen int            // float64 elements in host
	sum     uint64         // content hash of the flattened attributes the device copy holds
	stale   bool           // InvalidateMesh: upload on the next draw whatever the hash says
	lastUse uint64         // LRU stamp (Context.useClock)
}

// bytes the device copy occupies (planes + the staging area of the same size)
func (m *deviceMesh) deviceBytes() int { return 2 * 8 * m.hostLen }

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

// divCheck runs the device's self-check of its branch-free float64 division
// (fgl_debug_div_check): `pairs` generated operand pairs are divided with the
// helpers of the fused front end and with the operator; mismatches must be 0.
func (d *deviceContext) divCheck(seed, pairs uint64) (mismatches, fastPath uint64, err error) {
	var bad, fast C.uint64_t
	err = lastError(d.h, C.fgl_debug_div_check(d.h, C.uint64_t(seed), C.uint64_t(pairs), &bad, &fast))
	return uint64(bad), uint64(fast), err
}

// newTexture uploads an ImageTexture.  *image.RGBA (what Go's PNG decoder yields
// for 8-bit RGB) and *image.NRGBA (8-bit RGBA) are passed through as bytes;
// MakeColor's RGBA() conversion for each is reproduced on the device
// (color.go:25-29).  Every other image type -- *image.YCbCr from a JPEG
// (examples/capsule.go:35), Gray, Paletted, RGBA64, NRGBA64, CMYK ... -- is
// converted HERE with its own At(x, y).RGBA(), the call MakeColor makes, and
// uploaded as the resulting 16-bit values (FGL_TEX_RGBA64): exact for any
// image.Image, and the conversion stays in Go's image package.
func (d *deviceContext) newTexture(t *ImageTexture) (*deviceTexture, error) {
	var format int
	var bytes int
	w, h := t.Width, t.Height
	if w <= 0 || h <= 0 {
		return nil, errors.New("fauxgl: empty texture image")
	}
	var src []uint8
	var stride int
	format, bytes = C.FGL_TEX_RGBA64, 8*w*h
	switch im := t.Image.(type) { // (byte fast paths only for images whose origin is (0, 0), so that Pix row y is At(., y))
	case *image.RGBA:
		if im.Rect.Min == (image.Point{}) && im.Rect.Max.X >= w && im.Rect.Max.Y >= h {
			src, stride, format, bytes = im.Pix, im.Stride, C.FGL_TEX_RGBA, 4*w*h
		}
	case *image.NRGBA:
		if im.Rect.Min == (image.Point{}) && im.Rect.Max.X >= w && im.Rect.Max.Y >= h {
			src, stride, format, bytes = im.Pix, im.Stride, C.FGL_TEX_NRGBA, 4*w*h
		}
	}
	// stage in C memory: one copy, no Go pointers cross the boundary
	buf := C.malloc(C.size_t(bytes))
	if buf == nil {
		return nil, errors.New("fauxgl: out of memory staging a texture")
	}
	defer C.free(buf)
	if format == C.FGL_TEX_RGBA64 {
		dst := unsafe.Slice((*uint16)(buf), 4*w*h)
		i := 0
		for y := 0; y < h; y++ { // ImageTexture samples At(x, y) with x, y from 0 (texture.go:56-59)
			for x := 0; x < w; x++ {
				r, g, bl, a := t.Image.At(x, y).RGBA()
				dst[i], dst[i+1], dst[i+2], dst[i+3] = uint16(r), uint16(g), uint16(bl), uint16(a)
				i += 4
			}
		}
	} else {
		dst := unsafe.Slice((*uint8)(buf), bytes)
		for y := 0; y < h; y++ {
			copy(dst[4*w*y:4*w*(y+1)], src[stride*y:stride*y+4*w])
		}
	}
	var th *C.fgl_tex
	rc := C.fgl_texture_create(d.h, (*C.uint8_t)(buf), C.int(w), C.int(h), C.int(format), &th)
	if rc != 0 {
		return nil, lastError(d.h, rc)
	}
	return &deviceTexture{th}, nil
}

func (t *deviceTexture) destroy() { C.fgl_texture_destroy(t.h) }

// Layout of a device mesh's host block, in float64 elements: triangles [pos 9T | nrm 9T | tex 9T | col 12T],
// then lines [pos 6L | nrm 6L | tex 6L | col 8L] -- the [n][verts][k] arrays fgl_mesh_desc takes.
func hostElems(nt, nl int) int { return 39*nt + 26*nl }

// mix folds one float64 into a running 64-bit content hash (xor-multiply-rotate; not cryptographic: it only has
// to notice that a vertex changed between two draws of the same *Mesh).
func mix(h uint64, v float64) uint64 {
	h ^= math.Float64bits(v)
	h *= 0x9E3779B97F4A7C15
	return h<<29 | h>>35
}

// flattenInto gathers []*Triangle / []*Line into buf (C memory) and returns the content hash of what it wrote.
func flattenInto(mesh *Mesh, buf []float64) uint64 {
	nt, nl := len(mesh.Triangles), len(mesh.Lines)
	pos, nrm, tex, col := buf[0:9*nt], buf[9*nt:18*nt], buf[18*nt:27*nt], buf[27*nt:39*nt]
	h := uint64(nt)<<32 ^ uint64(nl)
	put := func(v *Vertex, i int, pos, nrm, tex, col []float64) {
		pos[3*i], pos[3*i+1], pos[3*i+2] = v.Position.X, v.Position.Y, v.Position.Z
		nrm[3*i], nrm[3*i+1], nrm[3*i+2] = v.Normal.X, v.Normal.Y, v.Normal.Z
		tex[3*i], tex[3*i+1], tex[3*i+2] = v.Texture.X, v.Texture.Y, v.Texture.Z
		col[4*i], col[4*i+1], col[4*i+2], col[4*i+3] = v.Color.R, v.Color.G, v.Color.B, v.Color.A
		h = mix(mix(mix(h, v.Position.X), v.Position.Y), v.Position.Z)
		h = mix(mix(mix(h, v.Normal.X), v.Normal.Y), v.Normal.Z)
		h = mix(mix(mix(h, v.Texture.X), v.Texture.Y), v.Texture.Z)
		h = mix(mix(mix(mix(h, v.Color.R), v.Color.G), v.Color.B), v.Color.A)
	}
	for i, t := range mesh.Triangles {
		put(&t.V1, 3*i, pos, nrm, tex, col)
		put(&t.V2, 3*i+1, pos, nrm, tex, col)
		put(&t.V3, 3*i+2, pos, nrm, tex, col)
	}
	l := buf[39*nt:]
	lpos, lnrm, ltex, lcol := l[0:6*nl], l[6*nl:12*nl], l[12*nl:18*nl], l[18*nl:26*nl]
	for i, ln := range mesh.Lines {
		put(&ln.V1, 2*i, lpos, lnrm, ltex, lcol)
		put(&ln.V2, 2*i+1, lpos, lnrm, ltex, lcol)
	}
	return h
}

// desc describes the host block to the library.  Every pointer in it is a C pointer (into m.host), so passing
// &desc -- itself Go memory -- obeys the cgo rules.
func (m *deviceMesh) desc() C.fgl_mesh_desc {
	var d C.fgl_mesh_desc
	nt, nl := m.nt, m.nl
	at := func(off int) *C.double {
		return (*C.double)(unsafe.Add(m.host, 8*off))
	}
	d.ntriangles, d.nlines = C.uint64_t(nt), C.uint64_t(nl)
	if nt > 0 {
		d.position, d.normal, d.texture, d.color = at(0), at(9*nt), at(18*nt), at(27*nt)
	}
	if nl > 0 {
		b := 39 * nt
		d.lposition, d.lnormal, d.ltexture, d.lcolor = at(b), at(b+6*nl), at(b+12*nl), at(b+18*nl)
	}
	return d
}

func (m *deviceMesh) hostSlice() []float64 {
	return unsafe.Slice((*float64)(m.host), m.hostLen)
}

func (d *deviceContext) newMesh(mesh *Mesh) (*deviceMesh, error) {
	m := &deviceMesh{nt: len(mesh.Triangles), nl: len(mesh.Lines)}
	m.hostLen = hostElems(m.nt, m.nl)
	if rc := C.fgl_host_alloc(C.size_t(8*m.hostLen), &m.host); rc != 0 {
		return nil, lastError(nil, rc)
	}
	m.sum = flattenInto(mesh, m.hostSlice())
	desc := m.desc()
	if rc := C.fgl_mesh_create(d.h, &desc, &m.h); rc != 0 {
		err := lastError(d.h, rc)
		C.fgl_host_free(m.host)
		return nil, err
	}
	return m, nil
}

func (m *deviceMesh) sameShape(mesh *Mesh) bool {
	return m.nt == len(mesh.Triangles) && m.nl == len(mesh.Lines)
}

// refresh re-flattens the host mesh and uploads it only if its content differs from what the device holds (or the
// mesh was invalidated).  Reports whether an upload happened.
func (m *deviceMesh) refresh(d *deviceContext, mesh *Mesh) (bool, error) {
	sum := flattenInto(mesh, m.hostSlice())
	if sum == m.sum && !m.stale {
		return false, nil
	}
	desc := m.desc()
	if rc := C.fgl_mesh_update(d.h, m.h, &desc); rc != 0 {
		return false, lastError(d.h, rc)
	}
	m.sum, m.stale = sum, false
	return true, nil
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
	if m.host != nil {
		C.fgl_host_free(m.host)
		m.host = nil
	}
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
	st.x_guard = cbool(s.xGuard)
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
