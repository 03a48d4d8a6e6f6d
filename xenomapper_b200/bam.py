"""BAM input for the walks: the replacement of the reference's `samtools view` pipes (xm.py:48-93).

Nothing is decoded in Python: the BGZF blocks are inflated, the record chain is followed and the alignment records are
rendered as SAM text by CUDA kernels (csrc/xm_inflate.h, xm_bamchain.h, xm_bam.h).
"""
import mmap
import os
import stat

from . import _lib


def _all_bytes(bamfile):
    """the file's bytes: a read-only mapping for a regular file (the library reads it once, front to back; nothing is
    copied into Python), else what read() returns"""
    try:
        fd = bamfile.fileno()
        st = os.fstat(fd)
        if stat.S_ISREG(st.st_mode) and st.st_size > 0 and "b" in getattr(bamfile, "mode", "b"):
            import numpy as np
            return np.frombuffer(mmap.mmap(fd, 0, access=mmap.ACCESS_READ), dtype=np.uint8)
    except (AttributeError, OSError, ValueError):
        pass
    bamfile.seek(0)
    data = bamfile.read()
    if isinstance(data, str):
        raise TypeError("BAM inputs must be opened in binary mode")
    return data


def read_header_text(bamfile):
    """the header text stored in the file (`samtools view -H`, xm.py:49); leaves the file at its start"""
    text = _lib.bam_header_text(_all_bytes(bamfile)).decode('ascii')
    bamfile.seek(0)
    return text


def records_as_sam_text(bamfile):
    """every alignment record as the SAM line `samtools view` prints (xm.py:61-64), as bytes"""
    return _lib.default_context().bam_render_host(_all_bytes(bamfile))
