// +build mksnap

package mksnap

import (
	"crypto/sha256"
	"encoding/hex"
	"fmt"
	"hash"
	"io"
	"io/ioutil"
	"os"
)

// TarAndGzip is the GPU counterpart of tarAndGzipDiffs (lib/builder/step/common.go:35-63): `commit` packs the layer
// (CommitCopyOps / CommitScan with tarOut = the write end of a pipe); the uncompressed tar bytes flow through
// `gzipTo` (tario.NewGzipWriter over ConcurrentMultiWriter{tempfile, gzipDigester}, unchanged Go) while the GPU
// digests the same arenas.  It returns the hex TarDigest (from the device) next to the gzip digester the caller
// already owns, i.e. exactly the two values commitLayer hex-encodes at common.go:86-87.
func TarAndGzip(sandboxDir string, newGzipWriter func(io.Writer) (io.WriteCloser, error),
	commit func(tarOut *os.File) (LayerResult, error)) (gzipDigester hash.Hash, tarHex string, name string, err error) {

	tmp, err := ioutil.TempFile(sandboxDir, "layertar-")
	if err != nil {
		return nil, "", "", fmt.Errorf("temp gzip tar file: %s", err)
	}
	defer tmp.Close()
	gzipDigester = sha256.New()
	gz, err := newGzipWriter(io.MultiWriter(tmp, gzipDigester))
	if err != nil {
		return nil, "", "", fmt.Errorf("new gzip writer: %s", err)
	}
	pr, pw, err := os.Pipe()
	if err != nil {
		return nil, "", "", fmt.Errorf("pipe: %s", err)
	}
	done := make(chan error, 1)
	go func() { // the gzip side drains the tar bytes while the packer fills the next arena
		_, cerr := io.Copy(gz, pr)
		if cerr == nil {
			cerr = gz.Close()
		}
		pr.Close()
		done <- cerr
	}()
	res, cerr := commit(pw)
	pw.Close()
	if gerr := <-done; cerr == nil && gerr != nil {
		cerr = fmt.Errorf("gzip layer: %s", gerr)
	}
	if cerr != nil {
		os.Remove(tmp.Name())
		return nil, "", "", fmt.Errorf("write diffs: %s", cerr)
	}
	return gzipDigester, res.TarDigest[len("sha256:"):], tmp.Name(), nil
}

// HexRoot is the chunk-table content address as stored under "<cache key>_chunks" (mkhost_cache_chunk_entry_create).
func (r LayerResult) HexRoot() string { return hex.EncodeToString(r.ChunkRoot[:]) }
