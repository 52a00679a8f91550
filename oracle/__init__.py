"""CPU ORACLE for the makisu snapshot+hash hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (makisu_b200, libmksnap.so) never does.

  oracle.lib       ctypes face of oracle/mkoracle.c (CRC-32, SHA-256, Roll-32 CDC, table root)
  oracle.ctx_crc   restatement of reference lib/builder/step/add_copy_step.go:102-238 (context cacheID)
  oracle.layer_tar restatement of reference lib/snapshot/mem_fs.go + mem_layer.go + lib/tario/write.go
                   and of Go 1.14 archive/tar's header writer (TarDigest byte stream)

Parity status:
  CRC-32 / SHA-256 arithmetic: pinned by the reference's own fixtures (tests/test_oracle_golden.py)
  USTAR header field formats : pinned by the Go-written fixture testdata/files/busybox/.../layer.tar
  cacheID byte order         : NO golden value in the reference -> "parity unpinned", pinned by us
  Roll-32 CDC / chunk table  : no reference counterpart        -> "parity unpinned", pinned by us
"""
from . import lib  # noqa: F401
