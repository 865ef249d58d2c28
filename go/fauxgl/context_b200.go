// Package fauxgl: B200 back end for the render context.
//
// This file REPLACES context.go of github.com/fogleman/fauxgl.  Every other
// file of the reference package (vector.go, matrix.go, color.go, vertex.go,
// triangle.go, line.go, mesh.go, shader.go, texture.go, loaders, shapes ...)
// stays as it is: they are host-side value types and mesh preparation, not the
// hot path.  The exported surface below is the reference's, name for name
// (context.go:11-58, 60, 83, 87, 119-145, 351-439); what changes is what runs
// behind it: hand-written sm_100a CUDA kernels reached through the C ABI of
// include/fauxgl_b200.h (see b200.go for the cgo binding and INTEGRATION.md).
//
// Differences a caller can observe, all documented in INTEGRATION.md:
//   - ColorBuffer / DepthBuffer live on the device between draws.  Image(),
//     DepthImage() and SyncBuffers() read them back into the exported fields;
//     a program that writes those fields itself must call UploadBuffers().
//   - A Shader (or Texture) that is not one of the built-ins has no device
//     equivalent: the draw returns a zero RasterizeInfo and Err() reports it.
//     There is no CPU fallback.
//   - Meshes are cached on the device by *Mesh pointer.  By default every draw
//     re-flattens the host mesh into the mesh's pinned C block and compares a
//     64-bit content hash over ALL vertex data with the device copy's: any
//     mutation -- Mesh methods or direct field pokes -- is caught and
//     re-uploaded; an unchanged mesh costs the flatten but no PCIe traffic.
//     A program that never mutates a mesh after its first draw (or calls
//     ctx.InvalidateMesh(mesh) when it does) sets ctx.MeshCache = MeshStatic
//     and pays nothing per draw.  The cache is bounded (MaxCachedMeshes
//     entries / MaxCachedMeshBytes of device memory, least recently used
//     evicted), so a program that builds a new Mesh per frame does not leak.
//   - Primitives are drawn in index order (the reference's goroutine schedule
//     is non-deterministic on depth ties, DepthBias, blending and
//     UpdatedPixels; index order is one of its legal schedules).
//
// NOTE: neither the build image nor the GPU box has a Go toolchain (probe:
// profiles/r02_go_probe.txt), so this file has been written to compile against
// the reference package but has not been compiled; all logic lives below the C
// ABI, which is tested through the ctypes mirror (fauxgl_b200/context.py) and
// the C++ mirror (include/fauxgl.hpp) that make exactly these calls.
package fauxgl

import (
	"errors"
	"image"
	"math"
	"runtime"
	"sync"
)

type Face int

const (
	_ Face = iota
	FaceCW
	FaceCCW
)

type Cull int

const (
	_ Cull = iota
	CullNone
	CullFront
	CullBack
)

type RasterizeInfo struct {
	TotalPixels   uint64
	UpdatedPixels uint64
}

func (info RasterizeInfo) Add(other RasterizeInfo) RasterizeInfo {
	return RasterizeInfo{
		info.TotalPixels + other.TotalPixels,
		info.UpdatedPixels + other.UpdatedPixels,
	}
}

type Context struct {
	Width       int
	Height      int
	ColorBuffer *image.NRGBA
	DepthBuffer []float64
	ClearColor  Color
	Shader      Shader
	ReadDepth   bool
	WriteDepth  bool
	WriteColor  bool
	AlphaBlend  bool
	Wireframe   bool
	FrontFace   Face
	Cull        Cull
	LineWidth   float64
	DepthBias   float64

	// Not in the reference.  XGuard: drop fragments with x outside [0, Width)
	// instead of letting them alias into the neighbouring row as context.go:223-228
	// does (the default, false, is the reference's behaviour).  MeshCache: see the
	// package comment.
	XGuard    bool
	MeshCache MeshCachePolicy

	// device side
	mu       sync.Mutex
	dev      *deviceContext        // b200.go
	meshes   map[*Mesh]*deviceMesh // device-resident copies, bounded LRU
	textures map[Texture]*deviceTexture
	useClock uint64
	err      error // sticky: first error of any call
	hostOK   bool  // exported buffers mirror the device
}

// MeshCachePolicy decides how DrawMesh notices that a cached *Mesh changed on the host.
type MeshCachePolicy int

const (
	// MeshCheck (default): re-flatten on every draw and compare a content hash of
	// all vertex data; upload only when it differs.  Never renders stale geometry.
	MeshCheck MeshCachePolicy = iota
	// MeshStatic: trust the device copy until InvalidateMesh(mesh) is called.
	MeshStatic
)

// Bounds of the device mesh cache (least recently used entries are evicted).
var (
	MaxCachedMeshes    = 64
	MaxCachedMeshBytes = 8 << 30
	MaxCachedTextures  = 32
)

// NewContext allocates the buffers on the device (GPU 0 unless FAUXGL_DEVICE
// selects another).  It panics only if no CUDA device exists, because the
// reference's signature has no error return and there is no CPU fallback.
func NewContext(width, height int) *Context {
	dc := &Context{}
	dc.Width = width
	dc.Height = height
	dc.ColorBuffer = image.NewNRGBA(image.Rect(0, 0, width, height))
	dc.DepthBuffer = make([]float64, width*height)
	dc.ClearColor = Transparent
	dc.Shader = NewSolidColorShader(Identity(), Color{1, 0, 1, 1})
	dc.ReadDepth = true
	dc.WriteDepth = true
	dc.WriteColor = true
	dc.AlphaBlend = true
	dc.Wireframe = false
	dc.FrontFace = FaceCCW
	dc.Cull = CullBack
	dc.LineWidth = 2
	dc.DepthBias = 0
	dc.meshes = make(map[*Mesh]*deviceMesh)
	dc.textures = make(map[Texture]*deviceTexture)
	dev, err := newDeviceContext(width, height, deviceFromEnv())
	if err != nil {
		panic("fauxgl: " + err.Error())
	}
	dc.dev = dev
	for i := range dc.DepthBuffer {
		dc.DepthBuffer[i] = math.MaxFloat64
	}
	dc.hostOK = true
	runtime.SetFinalizer(dc, (*Context).Close)
	return dc
}

// Close releases the device resources.  Safe to call more than once.
func (dc *Context) Close() {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	for _, m := range dc.meshes {
		m.destroy()
	}
	for _, t := range dc.textures {
		t.destroy()
	}
	dc.meshes, dc.textures = nil, nil
	if dc.dev != nil {
		dc.dev.destroy()
		dc.dev = nil
	}
}

// Err returns the first error any call on this context hit (unsupported
// shader, CUDA failure ...).  The reference's draw calls have no error return.
func (dc *Context) Err() error {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	return dc.err
}

func (dc *Context) fail(err error) {
	if dc.err == nil && err != nil {
		dc.err = err
	}
}

// SyncBuffers reads the device buffers back into ColorBuffer and DepthBuffer.
func (dc *Context) SyncBuffers() {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	dc.syncLocked()
}

func (dc *Context) syncLocked() {
	if dc.hostOK || dc.dev == nil {
		return
	}
	dc.fail(dc.dev.readColor(dc.ColorBuffer.Pix, dc.ColorBuffer.Stride))
	dc.fail(dc.dev.readDepth(dc.DepthBuffer))
	dc.hostOK = true
}

// UploadBuffers copies ColorBuffer and DepthBuffer to the device, for programs
// that write the exported buffers themselves.
func (dc *Context) UploadBuffers() {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	dc.fail(dc.dev.writeColor(dc.ColorBuffer.Pix, dc.ColorBuffer.Stride))
	dc.fail(dc.dev.writeDepth(dc.DepthBuffer))
	dc.hostOK = true
}

func (dc *Context) Image() image.Image {
	dc.SyncBuffers()
	return dc.ColorBuffer
}

func (dc *Context) DepthImage() image.Image {
	// normalised on the device (fgl_depth_image); only 2 bytes per pixel come back
	dc.mu.Lock()
	defer dc.mu.Unlock()
	im := image.NewGray16(image.Rect(0, 0, dc.Width, dc.Height))
	if dc.dev == nil {
		return im
	}
	{
		gray := make([]uint16, dc.Width*dc.Height)
		if err := dc.dev.depthImage(gray); err != nil {
			dc.fail(err)
			return im
		}
		for i, g := range gray { // Gray16.Pix is big-endian
			im.Pix[2*i] = uint8(g >> 8)
			im.Pix[2*i+1] = uint8(g)
		}
	}
	return im
}

// Resolve is the device version of resize.Resize(w/factor, h/factor, Image(),
// resize.Bilinear) that every example calls for supersampling: 16x fewer bytes
// cross PCIe than resizing on the host.  The result is premultiplied RGBA, as
// nfnt/resize returns for an *image.NRGBA input.
func (dc *Context) Resolve(factor int) (*image.RGBA, error) {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	out := image.NewRGBA(image.Rect(0, 0, dc.Width/factor, dc.Height/factor))
	err := dc.dev.resolve(factor, out.Pix)
	dc.fail(err)
	return out, err
}

func (dc *Context) ClearColorBufferWith(color Color) {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	c := color.NRGBA()
	dc.fail(dc.dev.clearColor(c.R, c.G, c.B, c.A))
	dc.hostOK = false
}

func (dc *Context) ClearColorBuffer() {
	dc.ClearColorBufferWith(dc.ClearColor)
}

func (dc *Context) ClearDepthBufferWith(value float64) {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	dc.fail(dc.dev.clearDepth(value))
	dc.hostOK = false
}

func (dc *Context) ClearDepthBuffer() {
	dc.ClearDepthBufferWith(math.MaxFloat64)
}

// InvalidateMesh forces a re-upload of the mesh on its next draw.
func (dc *Context) InvalidateMesh(mesh *Mesh) {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	if m, ok := dc.meshes[mesh]; ok {
		m.stale = true
	}
}

var errUnsupportedShader = errors.New(
	"fauxgl: Shader has no device implementation (only *SolidColorShader, *TextureShader, *PhongShader); there is no CPU fallback")

// describeShader resolves the Shader interface by type switch to the closed set
// with device implementations.
func (dc *Context) describeShader() (shaderDesc, error) {
	var d shaderDesc
	tex := func(t Texture) (*deviceTexture, error) {
		if t == nil {
			return nil, nil
		}
		if dt, ok := dc.textures[t]; ok {
			return dt, nil
		}
		it, ok := t.(*ImageTexture)
		if !ok {
			return nil, errors.New("fauxgl: Texture has no device implementation (only *ImageTexture)")
		}
		dt, err := dc.dev.newTexture(it)
		if err == nil {
			if len(dc.textures) >= MaxCachedTextures { // bounded: drop them all, they are re-uploaded on use
				for k, old := range dc.textures {
					old.destroy()
					delete(dc.textures, k)
				}
			}
			dc.textures[t] = dt
		}
		return dt, err
	}
	var err error
	switch s := dc.Shader.(type) {
	case *SolidColorShader:
		d.kind, d.matrix, d.color = shaderSolid, s.Matrix, s.Color
	case *TextureShader:
		d.kind, d.matrix = shaderTexture, s.Matrix
		d.texture, err = tex(s.Texture)
	case *PhongShader:
		d.kind, d.matrix = shaderPhong, s.Matrix
		d.light, d.camera = s.LightDirection, s.CameraPosition
		d.object, d.ambient, d.diffuse, d.specular = s.ObjectColor, s.AmbientColor, s.DiffuseColor, s.SpecularColor
		d.specularPower = s.SpecularPower
		d.texture, err = tex(s.Texture)
	default:
		err = errUnsupportedShader
	}
	return d, err
}

func (dc *Context) state() stateDesc {
	return stateDesc{dc.ReadDepth, dc.WriteDepth, dc.WriteColor, dc.AlphaBlend, dc.Wireframe,
		int(dc.FrontFace), int(dc.Cull), dc.LineWidth, dc.DepthBias, dc.XGuard}
}

// DrawTriangles and DrawLines take the reference's slices.  When the slice is
// exactly mesh.Triangles / mesh.Lines of a mesh drawn through DrawMesh the
// cached device copy is used; a bare slice is uploaded as a temporary mesh.
func (dc *Context) DrawTriangles(triangles []*Triangle) RasterizeInfo {
	return dc.draw(&Mesh{Triangles: triangles}, true, false, true)
}

func (dc *Context) DrawLines(lines []*Line) RasterizeInfo {
	return dc.draw(&Mesh{Lines: lines}, false, true, true)
}

func (dc *Context) DrawTriangle(t *Triangle) RasterizeInfo {
	return dc.DrawTriangles([]*Triangle{t})
}

func (dc *Context) DrawLine(l *Line) RasterizeInfo {
	return dc.DrawLines([]*Line{l})
}

func (dc *Context) DrawMesh(mesh *Mesh) RasterizeInfo {
	return dc.draw(mesh, true, true, false)
}

// DrawLinesEach returns what a loop of DrawLine over lines would return, one
// RasterizeInfo per line in index order, from a single launch sequence
// (examples/silhouette.go:163-166 tests each line's UpdatedPixels/TotalPixels).
func (dc *Context) DrawLinesEach(lines []*Line) []RasterizeInfo {
	return dc.drawEach(&Mesh{Lines: lines}, true)
}

// DrawTrianglesEach is the same for DrawTriangle.
func (dc *Context) DrawTrianglesEach(triangles []*Triangle) []RasterizeInfo {
	return dc.drawEach(&Mesh{Triangles: triangles}, false)
}

func (dc *Context) drawEach(mesh *Mesh, lines bool) []RasterizeInfo {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	n := len(mesh.Triangles)
	if lines {
		n = len(mesh.Lines)
	}
	result := make([]RasterizeInfo, n)
	if dc.dev == nil || n == 0 {
		return result
	}
	sh, err := dc.describeShader()
	if err != nil {
		dc.fail(err)
		return result
	}
	dm, err := dc.dev.newMesh(mesh)
	if err != nil {
		dc.fail(err)
		return result
	}
	defer dm.destroy()
	dc.fail(dc.dev.drawEach(dc.state(), sh, dm, lines, result))
	dc.hostOK = false
	return result
}

func (dc *Context) draw(mesh *Mesh, tris, lines, temporary bool) RasterizeInfo {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	var result RasterizeInfo
	if dc.dev == nil || (len(mesh.Triangles) == 0 && len(mesh.Lines) == 0) {
		return result
	}
	sh, err := dc.describeShader()
	if err != nil {
		dc.fail(err)
		return result
	}
	dm, err := dc.deviceMeshFor(mesh, temporary)
	if err != nil {
		dc.fail(err)
		return result
	}
	if temporary {
		defer dm.destroy()
	}
	st := dc.state()
	if tris && len(mesh.Triangles) > 0 {
		info, err := dc.dev.drawTriangles(st, sh, dm, 0, uint64(len(mesh.Triangles)))
		dc.fail(err)
		result = result.Add(info)
	}
	if lines && len(mesh.Lines) > 0 {
		info, err := dc.dev.drawLines(st, sh, dm, 0, uint64(len(mesh.Lines)))
		dc.fail(err)
		result = result.Add(info)
	}
	dc.hostOK = false
	return result
}

// deviceMeshFor returns the device copy of mesh, (re)uploading when the mesh is
// new, was invalidated, changed shape, or -- under MeshCheck -- its content hash
// differs from the device copy's.
func (dc *Context) deviceMeshFor(mesh *Mesh, temporary bool) (*deviceMesh, error) {
	if temporary {
		return dc.dev.newMesh(mesh)
	}
	dc.useClock++
	if m, ok := dc.meshes[mesh]; ok {
		if m.sameShape(mesh) {
			m.lastUse = dc.useClock
			if dc.MeshCache == MeshStatic && !m.stale {
				return m, nil
			}
			_, err := m.refresh(dc.dev, mesh)
			return m, err
		}
		m.destroy()
		delete(dc.meshes, mesh)
	}
	m, err := dc.dev.newMesh(mesh)
	if err != nil {
		return nil, err
	}
	m.lastUse = dc.useClock
	dc.meshes[mesh] = m
	dc.evictMeshes(m)
	return m, nil
}

// evictMeshes keeps the cache inside MaxCachedMeshes / MaxCachedMeshBytes, dropping
// the least recently used entries (never `keep`, the mesh about to be drawn).
func (dc *Context) evictMeshes(keep *deviceMesh) {
	for {
		total, n := 0, 0
		var oldestKey *Mesh
		var oldest *deviceMesh
		for k, m := range dc.meshes {
			total += m.deviceBytes()
			n++
			if m != keep && (oldest == nil || m.lastUse < oldest.lastUse) {
				oldestKey, oldest = k, m
			}
		}
		if oldest == nil || (n <= MaxCachedMeshes && total <= MaxCachedMeshBytes) {
			return
		}
		oldest.destroy()
		delete(dc.meshes, oldestKey)
	}
}
