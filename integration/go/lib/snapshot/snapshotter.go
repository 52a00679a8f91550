// Snapshotter is the interface north_star names; the reference only has the concrete *MemFS
// (lib/snapshot/mem_fs.go:60-83).  This file extracts its method set verbatim so that
// context.BuildContext.MemFS (lib/context/build_context.go:47) can hold either implementation.
// Place it in lib/snapshot next to mem_fs.go (no build tag: *MemFS satisfies it as is).
package snapshot

import "archive/tar"

// Snapshotter is the set of *MemFS methods the builder calls.
type Snapshotter interface {
	AddLayerByScan(w *tar.Writer) error
	AddLayerByCopyOps(cs []*CopyOperation, w *tar.Writer) error
	UpdateFromTarReader(r *tar.Reader, untar bool) error
	UpdateFromTarPath(source string, untar bool) error
	Checkpoint(newRoot string, sources []string) error
	Remove() error
	Reset()
}

var _ Snapshotter = (*MemFS)(nil)
