"""
xenomapper_b200.sharded -- the read-binning walk across the GPUs of one box.

One process per GPU (torchrun); `torch.distributed` carries a few dozen words
per rank and no SAM bytes.  The reference's walk (xm.py:95-118, 291-556) is a
single sequential pass, but record i depends only on record i-1 (pair
predicate xm.py:402, run-skipping reader xm.py:110-114) and on where the walk
stops (xm.py:105), so it shards by RECORD INDEX of the yielded sequence:

  1. every rank takes a line-aligned byte range of each input and runs the
     index pass on it (xm_count_device): records yielded, blank-line stop;
  2. all_gather of those counts -> every rank knows the record index at which
     each byte range starts in each stream and N = min(N1, N2);
  3. rank r owns records [N*r/W, N*(r+1)/W).  The owners of the byte ranges
     that hold the partition points locate them (xm_locate_device) and one
     all_reduce(MAX) spreads the byte offsets;
  4. each rank walks its record range of both streams with the record before
     it as context (XM_READER_FIRST_IS_CONTEXT), so pair units and QNAME runs
     that straddle a partition point are seen by exactly one rank;
  5. all_reduce(SUM) of the 36-bin histogram; all_gather of the six output
     lengths -> each rank's bytes go at the exclusive prefix: shard outputs
     concatenated in rank order equal the single-GPU output byte for byte.

The engine argument is what runs steps 1, 3 and 4 on the device; `GpuEngine`
wraps the C ABI.  (The test-suite passes the CPU emulation of the kernels to
cover this file's logic under gloo without a GPU.)
"""
import os

from . import _lib

_CHUNK = 1 << 16
_MAX = (1 << 64) - 1


# ---------------------------------------------------------------------------
# byte sources

class BytesSource:
    def __init__(self, data):
        self.data = bytes(data)
        self.size = len(self.data)

    def read(self, lo, hi):
        return self.data[lo:hi]


class FileSource:
    """the record region of a SAM file: everything from byte `offset` on"""

    def __init__(self, path_or_fd, offset=0):
        self.own = isinstance(path_or_fd, (str, bytes, os.PathLike))
        self.fd = os.open(path_or_fd, os.O_RDONLY) if self.own else path_or_fd
        self.offset = offset
        self.size = max(0, os.fstat(self.fd).st_size - offset)

    def read(self, lo, hi):
        out = bytearray()
        while lo < hi:
            b = os.pread(self.fd, min(hi - lo, 1 << 30), self.offset + lo)
            if not b:
                break
            out += b
            lo += len(b)
        return bytes(out)

    def close(self):
        if self.own:
            os.close(self.fd)


def line_start_at_or_after(src, x):
    """offset of the first line that starts at or after byte x"""
    if x <= 0:
        return 0
    p = x - 1                                   # a line starts at x iff byte x-1 is a newline
    while p < src.size:
        blk = src.read(p, min(src.size, p + _CHUNK))
        k = blk.find(b'\n')
        if k >= 0:
            return p + k + 1
        p += len(blk)
    return src.size


def previous_line_start(src, x):
    """offset of the line before the one that starts at x (x > 0 is a line start)"""
    end = x - 1                                 # the newline that ends the previous line
    while end > 0:
        lo = max(0, end - _CHUNK)
        blk = src.read(lo, end)
        k = blk.rfind(b'\n')
        if k >= 0:
            return lo + k + 1
        end = lo
    return 0


# ---------------------------------------------------------------------------
# engines

class GpuEngine:
    """index pass and walk of one rank through libxenomapper_b200.so"""

    def __init__(self, ctx=None):
        self.ctx = ctx or _lib.default_context()

    def index(self, buf, queries, skip):
        """(records, stopped, end offset, byte offsets of the queried records) of one resident byte range;
        a call with queries answers only those (the counts were exchanged by then)"""
        ctx = self.ctx
        if not buf:
            return 0, False, 0, [0] * len(queries)
        d = ctx.dev_alloc(len(buf))
        try:
            ctx.h2d(d, buf)
            if queries:
                return None, None, None, ctx.locate_device(d, len(buf), list(queries), skip)
            info = ctx.count_device(d, len(buf), skip)
            return int(info.n_records), info.stop_at != _MAX, int(info.end_off), []
        finally:
            ctx.dev_free(d)

    def walk(self, prim, sec, mode, score_src, skip, min_score, enabled_bins, first_is_context):
        ctx = self.ctx
        opts = ctx.opts(mode, score_src, skip, float(min_score), enabled_bins, first_is_context=first_is_context)
        rc, res, outs = ctx.classify_host(prim, sec, opts)
        return dict(status=rc, counts=list(res.counts), outputs=outs, n_records=int(res.n_records),
                    err_record=int(res.err_record), message=ctx.error() if rc else "")


# ---------------------------------------------------------------------------
# the collective plumbing: torch.distributed when there is more than one rank

class _Comm:
    def __init__(self, group=None, device=None):
        self.dist = None
        self.rank, self.size = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.dist = dist
                self.group = group
                self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        except ImportError:
            pass
        if self.dist is not None:
            import torch
            self.torch = torch
            if device is None:
                device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            self.device = device

    def all_gather(self, words):
        """list of int64 per rank -> list (by rank) of lists"""
        if self.dist is None:
            return [list(words)]
        t = self.torch.tensor(list(words), dtype=self.torch.int64, device=self.device)
        out = [self.torch.empty_like(t) for _ in range(self.size)]
        self.dist.all_gather(out, t, group=self.group)
        return [[int(v) for v in o.cpu().tolist()] for o in out]

    def all_reduce(self, words, op="sum"):
        if self.dist is None:
            return list(words)
        t = self.torch.tensor(list(words), dtype=self.torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX, group=self.group)
        return [int(v) for v in t.cpu().tolist()]


# ---------------------------------------------------------------------------

def plan_partition(n, world):
    """record-index partition points of n records over `world` ranks"""
    return [n * k // world for k in range(world + 1)]


def sharded_walk(engine, prim, sec, mode=_lib.MODE_SE, score_src=_lib.SCORE_AS_XS, skip=False, min_score=float("-inf"),
                 enabled_bins=0x3F, group=None, device=None):
    """Walk this rank's record-index shard of (prim, sec).  Returns a dict:
    status (first failing rank's, 0 if none), message, counts (36, whole job), n_records (whole job),
    outputs (this rank's six byte strings), out_offset (where they go inside each bin, counted from the end of
    the bin's header), out_total (six whole-job lengths), records (this rank's [lo, hi))."""
    comm = _Comm(group, device)
    W, r = comm.size, comm.rank
    srcs = (prim, sec)

    # 1. line-aligned byte ranges; with skip the line before the range is context for the run-skipping reader
    rng = []
    for s in srcs:
        lo = line_start_at_or_after(s, s.size * r // W)
        hi = line_start_at_or_after(s, s.size * (r + 1) // W) if r + 1 < W else s.size
        ctx_lo = previous_line_start(s, lo) if (skip and lo > 0) else lo
        rng.append((ctx_lo, lo, max(lo, hi)))
    bufs = [s.read(c, h) for s, (c, l, h) in zip(srcs, rng)]
    has_ctx = [c < l for (c, l, h) in rng]
    mine = []
    for k in range(2):
        n, stopped, end_off, _ = engine.index(bufs[k], (), skip)
        n = max(0, n - (1 if has_ctx[k] else 0)) if rng[k][1] < rng[k][2] else 0
        # a range that is only context yields nothing of its own
        if rng[k][1] >= rng[k][2]:
            stopped = False
        mine += [n, 1 if stopped else 0]

    # 2. record index at which every byte range starts, per stream; the stream ends at its first blank line
    allc = comm.all_gather(mine)
    base = [[0] * (W + 1), [0] * (W + 1)]
    alive_last = [W - 1, W - 1]
    for k in range(2):
        dead = False
        for q in range(W):
            n = 0 if dead else allc[q][2 * k]
            base[k][q + 1] = base[k][q] + n
            if not dead and allc[q][2 * k + 1]:
                dead = True
                alive_last[k] = q
    total = [base[0][W], base[1][W]]
    N = min(total)
    cuts = plan_partition(N, W)

    # 3. byte offsets of the partition points (and of the record before each: the context record)
    want = sorted(set(cuts) | {c - 1 for c in cuts[1:W] if c > 0})
    where = {}
    ans = []
    for k in range(2):
        q_local, q_glob = [], []
        for g in want:
            owner = base[k][r] <= g < base[k][r + 1] or (g == total[k] and r == alive_last[k])
            if owner:
                q_glob.append(g)
                q_local.append(g - base[k][r] + (1 if has_ctx[k] else 0))
        offs = engine.index(bufs[k], q_local, skip)[3] if q_local else []
        found = dict(zip(q_glob, (rng[k][0] + o for o in offs)))
        ans += [found.get(g, -1) for g in want]
    ans = comm.all_reduce(ans, op="max")
    for k in range(2):
        for j, g in enumerate(want):
            where[(k, g)] = ans[k * len(want) + j]
    del bufs

    # 4. this rank's record range, preceded by its context record
    lo_rec, hi_rec = cuts[r], cuts[r + 1]
    ctx = lo_rec > 0 and hi_rec > lo_rec
    res = dict(status=0, counts=[0] * 36, outputs=[b""] * 6, n_records=0, err_record=0, message="")
    if hi_rec > lo_rec:
        parts = []
        for k in range(2):
            a = where[(k, lo_rec - 1 if ctx else lo_rec)]
            b = where[(k, hi_rec)]
            if a < 0 or b < 0:
                raise RuntimeError("partition point not located (stream %d, records %d..%d)" % (k, lo_rec, hi_rec))
            parts.append(srcs[k].read(a, b))
        res = engine.walk(parts[0], parts[1], mode, score_src, skip, min_score, enabled_bins, ctx)
        del parts

    # 5. histogram, output placement, first failure
    lens = [len(o) for o in res["outputs"]]
    info = comm.all_gather([res["status"], lo_rec + res["err_record"], res["n_records"]] + lens)
    failed = [q for q in range(W) if info[q][0] != 0]
    first_bad = failed[0] if failed else W
    # ranks after a failing one contribute nothing: the reference stops at the failing record (streaming writes)
    if r > first_bad:
        res["outputs"] = [b""] * 6
        res["counts"] = [0] * 36
        lens = [0] * 6
    counts = comm.all_reduce(res["counts"], op="sum")
    out_offset = [sum(info[q][3 + b] for q in range(min(r, first_bad + 1))) for b in range(6)]
    out_total = [sum(info[q][3 + b] for q in range(min(W, first_bad + 1))) for b in range(6)]
    n_done = sum(info[q][2] for q in range(min(W, first_bad + 1)))
    status, message, err_record = 0, "", 0
    if failed:
        status, err_record = info[first_bad][0], info[first_bad][1]
        message = res["message"] if r == first_bad else "rank %d failed at record %d" % (first_bad, err_record)
    return dict(status=status, message=message, err_record=err_record, counts=counts, n_records=n_done,
                outputs=res["outputs"], out_offset=out_offset, out_total=out_total, records=(lo_rec, hi_rec),
                rank=r, world=W)


def write_outputs(result, fds, header_len):
    """pwrite this rank's bytes of each enabled bin behind the bin's header (fds[b] < 0: disabled)"""
    for b in range(6):
        if fds[b] is None or fds[b] < 0:
            continue
        data, at = result["outputs"][b], header_len[b] + result["out_offset"][b]
        done = 0
        while done < len(data):
            done += os.pwrite(fds[b], data[done:done + (1 << 30)], at + done)
