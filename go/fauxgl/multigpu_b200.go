// Sort-last rendering across the GPUs of one box from Go: the bindings of the
// composite entry points of include/fauxgl_b200.h (fgl_peer_* and fgl_comm_*).
// Not part of the reference's API -- the reference reduces its goroutines'
// results over a channel inside one address space (context.go:393-410); ranks
// on different GPUs exchange pixels instead.
//
// One Context per GPU (FAUXGL_DEVICE, or one process per GPU).  Every rank
// draws its share of the triangles with DrawMeshRange, then all ranks call
// Composite once per frame; afterwards the presenting rank (or every rank)
// holds the frame.  Valid for the order-independent state only: ReadDepth,
// WriteDepth, DepthBias 0, opaque output.
//
//	ctx := fauxgl.NewContext(w, h)                    // this rank's GPU
//	rec, _ := ctx.PeerExport()                        // exchange with the other ranks (channel, pipe, MPI ...)
//	group, _ := ctx.NewPeerGroup(rank, allRecords)    // records of all ranks, in rank order
//	for frame := range frames {
//		ctx.ClearDepthBuffer(); ctx.ClearColorBufferWith(bg)
//		ctx.DrawMeshRange(mesh, first, count)         // this rank's triangles
//		group.Composite(0)                            // rank 0 presents
//	}
//
// NOTE: like the rest of this package the file has not been compiled (no Go
// toolchain in the build image or on the GPU box: profiles/r02_go_probe.txt).
package fauxgl

/*
#include <stdlib.h>
#include "fauxgl_b200.h"
*/
import "C"

import (
	"errors"
	"unsafe"
)

// PeerExport returns this context's record for NewPeerGroup (CUDA IPC handles of
// its colour, depth, dirty-strip and flag buffers; C.FGL_PEER_EXPORT_BYTES).
func (dc *Context) PeerExport() ([]byte, error) {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	rec := make([]byte, C.FGL_PEER_EXPORT_BYTES)
	rc := C.fgl_peer_export(dc.dev.h, unsafe.Pointer(&rec[0]))
	return rec, lastError(dc.dev.h, rc)
}

// PeerGroup is the exact, sparse, fused composite over peer memory
// (fgl_peer_group): float64 depth, ties to the higher rank, only strips a rank
// has drawn are read from it, flags in peer memory instead of host barriers.
type PeerGroup struct {
	h  *C.fgl_peer_group
	dc *Context
}

// NewPeerGroup opens the peers' buffers.  records holds the PeerExport record of
// every rank, in rank order (this rank's own included).
func (dc *Context) NewPeerGroup(rank int, records [][]byte) (*PeerGroup, error) {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	n := len(records)
	if n == 0 || rank < 0 || rank >= n {
		return nil, errors.New("fauxgl: bad rank / empty record list")
	}
	blob := make([]byte, 0, n*int(C.FGL_PEER_EXPORT_BYTES))
	for _, r := range records {
		if len(r) != int(C.FGL_PEER_EXPORT_BYTES) {
			return nil, errors.New("fauxgl: peer record has the wrong size")
		}
		blob = append(blob, r...)
	}
	var h *C.fgl_peer_group
	rc := C.fgl_peer_group_create(dc.dev.h, C.int(rank), C.int(n), unsafe.Pointer(&blob[0]), &h)
	if rc != 0 {
		return nil, lastError(dc.dev.h, rc)
	}
	return &PeerGroup{h, dc}, nil
}

// Composite enqueues signal -> wait -> sparse composite -> signal -> wait on the
// context's stream and returns at once.  Afterwards rank root holds the frame
// (every rank if root < 0).  Collective: every rank calls it once per frame.
func (g *PeerGroup) Composite(root int) error { return g.composite(root, 0) }

// CompositeColors is Composite for a rank that only PRESENTS the frame: it receives
// the winning colours but not the depths (FGL_COMPOSITE_COLOR_ONLY), a third of the
// bytes; its depth buffer stays as it drew it.
func (g *PeerGroup) CompositeColors(root int) error {
	return g.composite(root, C.FGL_COMPOSITE_COLOR_ONLY)
}

func (g *PeerGroup) composite(root int, flags C.int) error {
	g.dc.mu.Lock()
	defer g.dc.mu.Unlock()
	g.dc.hostOK = false
	return lastError(g.dc.dev.h, C.fgl_peer_composite(g.dc.dev.h, g.h, C.int(root), flags))
}

// Status waits for the stream and reports a rank that never arrived.
func (g *PeerGroup) Status() error {
	g.dc.mu.Lock()
	defer g.dc.mu.Unlock()
	return lastError(g.dc.dev.h, C.fgl_peer_status(g.dc.dev.h, g.h))
}

func (g *PeerGroup) Close() {
	if g.h != nil {
		C.fgl_peer_group_destroy(g.h)
		g.h = nil
	}
}

// CommUniqueID creates the id rank 0 hands to the other ranks for NewComm
// (ncclGetUniqueId; NCCL is loaded by the library itself).
func CommUniqueID() ([]byte, error) {
	id := make([]byte, C.FGL_COMM_ID_BYTES)
	rc := C.fgl_comm_unique_id(unsafe.Pointer(&id[0]))
	return id, lastError(nil, rc)
}

// Comm is the packed-key composite over NCCL (fgl_comm): keys
// (depth32<<32 | rgba8) min-reduced by screen stripe with ncclReduceScatter,
// then gathered to the presenting rank or all-gathered.
type Comm struct {
	h  *C.fgl_comm
	dc *Context
}

// NewComm joins the communicator (collective: blocks until all ranks have).
func (dc *Context) NewComm(nranks, rank int, id []byte) (*Comm, error) {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	if len(id) != int(C.FGL_COMM_ID_BYTES) {
		return nil, errors.New("fauxgl: NCCL id has the wrong size")
	}
	var h *C.fgl_comm
	rc := C.fgl_comm_init(dc.dev.h, C.int(nranks), C.int(rank), unsafe.Pointer(&id[0]), &h)
	if rc != 0 {
		return nil, lastError(dc.dev.h, rc)
	}
	return &Comm{h, dc}, nil
}

func (c *Comm) Composite(root int) error {
	c.dc.mu.Lock()
	defer c.dc.mu.Unlock()
	c.dc.hostOK = false
	return lastError(c.dc.dev.h, C.fgl_composite(c.dc.dev.h, c.h, C.int(root)))
}

func (c *Comm) Close() {
	if c.h != nil {
		C.fgl_comm_destroy(c.h)
		c.h = nil
	}
}

// DrawMeshRange draws triangles [first, first+count) of the (cached) mesh: one
// rank's share of a sort-last frame.
func (dc *Context) DrawMeshRange(mesh *Mesh, first, count int) RasterizeInfo {
	dc.mu.Lock()
	defer dc.mu.Unlock()
	var result RasterizeInfo
	if dc.dev == nil || count <= 0 || first < 0 || first+count > len(mesh.Triangles) {
		return result
	}
	sh, err := dc.describeShader()
	if err != nil {
		dc.fail(err)
		return result
	}
	dm, err := dc.deviceMeshFor(mesh, false)
	if err != nil {
		dc.fail(err)
		return result
	}
	info, err := dc.dev.drawTriangles(dc.state(), sh, dm, uint64(first), uint64(count))
	dc.fail(err)
	dc.hostOK = false
	return info
}
