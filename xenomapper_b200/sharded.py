"""
xenomapper_b200.sharded -- the read-binning walk across the GPUs of one box.

One process per GPU, started by any launcher that sets RANK / WORLD_SIZE / LOCAL_RANK
(`python -m <launcher> --nproc-per-node N -m xenomapper_b200.xenomapper ...`).  The walk itself, its NCCL communicator
and every collective live inside libxenomapper_b200.so (csrc/xm_shard.h): this module only

  * ferries the 128-byte NCCL id from rank 0 to the other ranks (a file in the launch's
    scratch directory -- one box, one file system; no framework, no MPI),
  * gives every rank its BYTE shard of the two record regions, bytes [size*r/W, size*(r+1)/W),
    cut anywhere: the library moves the line heads, context lines and record slivers between
    neighbours and walks the RECORDS the rank's primary shard holds (the reference's lockstep
    reader pairs the files by record index, xm.py:95-118),
  * writes the rank's part of each bin in place behind the header: shard outputs concatenated
    in rank order equal the single-GPU output byte for byte.
"""
import os
import struct
import tempfile
import time

from . import _lib


# ---------------------------------------------------------------------------
# rendezvous: small blobs from rank 0 to everybody, through files of the launch's scratch directory

class Rendezvous:
    """publish(name, blob) on one rank, fetch(name) on the others.  Files live in a directory keyed by the launch
    (XM_COMM_DIR, else MASTER_PORT + the launcher's pid, which all ranks of one launch share)."""

    def __init__(self, rank, world, directory=None, timeout=120.0):
        self.rank, self.world, self.timeout = rank, world, timeout
        if directory is None:
            directory = os.environ.get("XM_COMM_DIR")
        if directory is None:
            key = "%s_%s_%d" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "none"), os.getppid())
            directory = os.path.join(tempfile.gettempdir(), "xm_comm_" + key)
        os.makedirs(directory, exist_ok=True)
        self.dir = directory

    def publish(self, name, blob):
        path = os.path.join(self.dir, name)
        tmp = "%s.%d.tmp" % (path, os.getpid())
        with open(tmp, "wb") as f:
            f.write(blob)
        os.replace(tmp, path)                     # atomic: a reader sees the whole blob or nothing

    def fetch(self, name):
        path = os.path.join(self.dir, name)
        t0 = time.monotonic()
        while True:
            try:
                with open(path, "rb") as f:
                    return f.read()
            except FileNotFoundError:
                if time.monotonic() - t0 > self.timeout:
                    raise TimeoutError("rank %d: nothing published as %s within %.0f s" % (self.rank, path, self.timeout))
                time.sleep(0.005)

    def cleanup(self):
        if self.rank == 0:
            for f in os.listdir(self.dir):
                try:
                    os.unlink(os.path.join(self.dir, f))
                except OSError:
                    pass
            try:
                os.rmdir(self.dir)
            except OSError:
                pass


def init_comm(ctx, rank, world, rendezvous=None, tag="nccl_id"):
    """the library's communicator for this context: rank 0 makes the NCCL id, the rendezvous carries it"""
    if world == 1:
        ctx.comm_init_rank(1, 0, None)
        return None
    rv = rendezvous or Rendezvous(rank, world)
    if rank == 0:
        rv.publish(tag, ctx.comm_unique_id())
    ctx.comm_init_rank(world, rank, rv.fetch(tag))
    return rv


# ---------------------------------------------------------------------------
# byte shards

def byte_range(size, rank, world):
    return size * rank // world, size * (rank + 1) // world


def read_range(fd, lo, hi):
    """bytes [lo, hi) of a descriptor"""
    out = bytearray(hi - lo)
    view, done = memoryview(out), 0
    while done < hi - lo:
        n = os.preadv(fd, [view[done:done + (1 << 30)]], lo + done)
        if n <= 0:
            break
        done += n
    return bytes(out[:done])


def sharded_walk(ctx, rank, world, prim_fd, prim_off, sec_fd, sec_off, mode=_lib.MODE_SE, score_src=_lib.SCORE_AS_XS,
                 skip=False, min_score=float("-inf"), enabled_bins=0x3F):
    """Walk this rank's shard of the record regions that start at byte prim_off / sec_off of the two descriptors.
    Returns a dict: status, message, counts (36, whole job), n_records (whole job), outputs (this rank's six byte
    strings), out_offset / out_total (where they go inside each bin), records (this rank's [lo, hi)), stats."""
    sizes = [max(0, os.fstat(fd).st_size - off) for fd, off in ((prim_fd, prim_off), (sec_fd, sec_off))]
    shards = []
    for fd, off, size in ((prim_fd, prim_off, sizes[0]), (sec_fd, sec_off, sizes[1])):
        lo, hi = byte_range(size, rank, world)
        shards.append(read_range(fd, off + lo, off + hi))
    opts = ctx.opts(mode, score_src, skip, float(min_score), enabled_bins)
    rc, res, st, outs = ctx.classify_sharded_host(shards[0], shards[1], opts)
    return dict(status=rc, message=ctx.error() if rc else "", err_record=int(res.err_record), counts=list(res.counts),
                n_records=int(res.n_records), outputs=outs, out_offset=list(st.out_offset), out_total=list(st.out_total),
                records=(int(st.rec_lo), int(st.rec_hi)), rank=rank, world=world,
                stats=dict(align_ms=st.align_ms, index_ms=st.index_ms, sliver_ms=st.sliver_ms, walk_ms=st.walk_ms,
                           comm_ms=st.comm_ms, total_ms=st.total_ms, sliver_bytes=int(st.sliver_bytes),
                           sent_bytes=int(st.sent_bytes), collectives=int(st.n_collectives)))


def write_outputs(result, fds, header_len):
    """pwrite this rank's bytes of each enabled bin behind the bin's header (fds[b] < 0: disabled)"""
    for b in range(6):
        if fds[b] is None or fds[b] < 0:
            continue
        data, at = result["outputs"][b], header_len[b] + result["out_offset"][b]
        done = 0
        while done < len(data):
            done += os.pwrite(fds[b], data[done:done + (1 << 30)], at + done)


# ---------------------------------------------------------------------------
# BAM input: every rank maps both files and takes the BGZF blocks that start in its 1/world of each file's bytes

def _settle_entries(ctx, rank, world, rv, stream, guess, exit_off, tag):
    """The ranks' record chains joined: rank r's part begins where the chain of the ranks before it leaves off.  Every rank
    publishes (entry it followed, exit it reached) -- (NONE, NONE) when it saw no record start of its own -- and all read the
    same table: going through the ranks in order, a rank whose entry is not the exit handed to it follows its part again from
    there (xm_bam_shard_chain: a part the chain jumps over answers with the entry itself) and the table is made anew.
    Guesses are nearly always right: one round, or two when a part saw nothing."""
    none = ctx.NONE64
    entry = guess
    for rnd in range(world + 2):
        rv.publish("%s_%d_%d_%d" % (tag, stream, rnd, rank), struct.pack("<QQ", entry, exit_off))
        rows = [struct.unpack("<QQ", rv.fetch("%s_%d_%d_%d" % (tag, stream, rnd, r))) for r in range(world)]
        cur, mine, ok = None, None, True
        for r, (e_used, ex) in enumerate(rows):
            if r == rank:
                mine = cur
            if cur is None:
                if e_used != none:
                    cur = ex                 # the first rank that holds records knows where they start
                continue
            if e_used != cur:
                ok = False                   # that rank looks again; what it reports now cannot be trusted
            cur = ex if e_used != none else cur
        if ok:
            return
        if mine is not None and entry != mine:
            exit_off = ctx.bam_shard_chain(stream, mine)
            entry = mine
    raise RuntimeError("the ranks' BAM record chains did not settle")


def sharded_bam_walk(ctx, rank, world, rv, prim_bam, sec_bam, mode=_lib.MODE_SE, score_src=_lib.SCORE_AS_XS, skip=False,
                     min_score=float("-inf"), enabled_bins=0x3F, room=1 << 20, tag="bamchain", want_outputs=True):
    """prim_bam / sec_bam: the WHOLE files (bytes-like; numpy arrays over mapped files).  This rank's part of each is
    inflated and rendered as SAM text on its GPU, the text shards go through the walk across GPUs.  Returns the dict of
    sharded_walk()."""
    if rv is None and world > 1:
        rv = Rendezvous(rank, world)
    shards = []
    for stream, bam in enumerate((prim_bam, sec_bam)):
        guess, exit_off = ctx.bam_shard_open(stream, bam, rank, world)
        if world > 1:
            _settle_entries(ctx, rank, world, rv, stream, guess, exit_off, tag)
        shards.append(ctx.bam_shard_text(stream, room, room))
    (dp, np_), (ds, ns_) = shards
    opts = ctx.opts(mode, score_src, skip, float(min_score), enabled_bins)
    slack = 2 * room + 4096
    mul = 1 if mode == _lib.MODE_SE else 2             # overlapping pair units can emit a line twice (the library's own bound)
    pb, sb = (np_ + slack) * mul, (ns_ + slack) * mul
    caps = [pb, sb, pb, sb, pb, pb + sb]               # PS SS PM SM UA UR
    caps = [c if (enabled_bins >> b) & 1 else 16 for b, c in enumerate(caps)]
    d_out = [ctx.dev_alloc(c) for c in caps]
    try:
        rc, res, st = ctx.classify_sharded_device(dp, np_, ds, ns_, room, room, opts, d_out, caps)
        outs = [ctx.d2h(d_out[b], int(res.out_len[b])) if rc == _lib.XM_OK and want_outputs else b"" for b in range(6)]
    finally:
        for d in d_out:
            ctx.dev_free(d)
    return dict(status=rc, message=ctx.error() if rc else "", err_record=int(res.err_record), counts=list(res.counts),
                n_records=int(res.n_records), outputs=outs, out_offset=list(st.out_offset), out_total=list(st.out_total),
                records=(int(st.rec_lo), int(st.rec_hi)), rank=rank, world=world,
                stats=dict(align_ms=st.align_ms, index_ms=st.index_ms, sliver_ms=st.sliver_ms, walk_ms=st.walk_ms,
                           comm_ms=st.comm_ms, total_ms=st.total_ms, sliver_bytes=int(st.sliver_bytes),
                           sent_bytes=int(st.sent_bytes), collectives=int(st.n_collectives)))
