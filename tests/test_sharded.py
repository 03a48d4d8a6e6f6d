"""The sharded walk (xenomapper_b200/sharded.py) under gloo with world sizes 2 and 3.

Every rank runs the emulated kernels on its record-index shard; the concatenated outputs and the reduced
histogram must equal the oracle's single-pass result byte for byte.  GPU twin: test_gpu_parity.py
(test_sharded_walk_on_gpu) runs the same driver through the C ABI on one device per rank.
"""
import json
import os
import socket
import subprocess
import sys

import pytest

from oracle import oracle
from xenomapper_b200 import sharded, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def with_repeats(buf, every=7):
    """duplicate every `every`-th line (multi-mapping style repeats for the run-skipping reader)"""
    out = []
    for i, line in enumerate(bytes(buf).split(b"\n")[:-1]):
        out.append(line)
        if i % every == 3:
            out.append(line)
            if i % (3 * every) == 3:
                out.append(line)
    return b"\n".join(out) + b"\n"


def make_case(kind):
    if kind == "se_skip":
        p, s = synth.generate(1500, seed=11, style=synth.STYLE_SE_BOWTIE2)
        return with_repeats(p), with_repeats(s, every=5), dict(mode=0, score_src=0, skip=True, min_score=float("-inf"))
    if kind == "se_noskip_short_secondary":
        p, s = synth.generate(1200, seed=12, style=synth.STYLE_SE_BOWTIE2)
        s = bytes(s)
        cut = s.rfind(b"\n", 0, len(s) * 2 // 3) + 1
        return bytes(p), s[:cut], dict(mode=0, score_src=0, skip=False, min_score=float("-inf"))
    if kind == "pe_liberal":
        p, s = synth.generate(1600, seed=13, style=synth.STYLE_PE_BOWTIE2)
        return bytes(p), bytes(s), dict(mode=1, score_src=0, skip=False, min_score=float("-inf"))
    if kind == "pe_conservative_zs":
        p, s = synth.generate(1600, seed=14, style=synth.STYLE_PE_HISAT)
        return bytes(p), bytes(s), dict(mode=2, score_src=1, skip=False, min_score=-18.0)
    if kind == "pe_cigar_blank_stop":
        p, s = synth.generate(1400, seed=15, style=synth.STYLE_PE_BOWTIE2)
        p = bytes(p)
        cut = p.rfind(b"\n", 0, len(p) * 3 // 5) + 1
        return p[:cut] + b"\n" + p[cut:], bytes(s), dict(mode=1, score_src=2, skip=False, min_score=-40.0)
    if kind == "tiny":
        p, s = synth.generate(3, seed=16, style=synth.STYLE_PE_BOWTIE2)
        return bytes(p), bytes(s), dict(mode=1, score_src=0, skip=False, min_score=float("-inf"))
    raise KeyError(kind)


def run_sharded(tmp_path, world, p, s, o, engine="emu", room=1 << 20):
    files = {}
    for name, data in (("prim", p), ("sec", s)):
        files[name] = str(tmp_path / (name + ".sam"))
        open(files[name], "wb").write(data)
    outs = [str(tmp_path / ("out%d.sam" % b)) for b in range(6)]
    for f in outs:
        open(f, "wb").close()
    case = dict(files, outs=outs, result=str(tmp_path / "result.json"), enabled_bins=0x3F, engine=engine, room=room,
                rendezvous=str(tmp_path / "rv"), **o)
    casefile = str(tmp_path / "case.json")
    json.dump(case, open(casefile, "w"))
    port = free_port()
    procs = [subprocess.Popen([sys.executable, "-m", "tests._sharded_worker", str(r), str(world), str(port), casefile],
                              cwd=ROOT, stderr=subprocess.PIPE) for r in range(world)]
    for pr in procs:
        _, err = pr.communicate(timeout=600)
        assert pr.returncode == 0, err.decode()[-2000:]
    per_rank = [json.load(open("%s.%d" % (case["result"], r))) for r in range(world)]
    return per_rank, [open(f, "rb").read() for f in outs]


def check_against_oracle(per_rank, outs, ref, world):
    """rank-order concatenated bins, the summed histogram and the record ranges equal the single pass"""
    first = per_rank[0]
    assert all(r["status"] == 0 for r in per_rank), [r["message"] for r in per_rank]
    if first.get("declined"):
        counts, n = first["counts"], first["n_records"]
    else:
        # every rank reports the whole job's histogram and record count
        assert all(r["counts"] == first["counts"] and r["n_records"] == first["n_records"] for r in per_rank)
        counts, n = first["counts"], first["n_records"]
        # the ranks' record ranges tile [0, n) in rank order (a rank may walk nothing: after a gather on rank 0 all but rank 0)
        spans = [tuple(r["records"]) for r in per_rank if r["records"][1] > r["records"][0]]
        assert (not spans and n == 0) or (spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:])))
    assert counts == ref["counts"]
    assert n == ref["n_yielded"]
    assert first["out_total"] == [len(x) for x in ref["outputs"]]
    assert outs == ref["outputs"]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("kind", ["se_skip", "se_noskip_short_secondary", "pe_liberal", "pe_conservative_zs",
                                  "pe_cigar_blank_stop", "tiny"])
def test_sharded_walk_equals_single_pass(tmp_path, kind, world):
    p, s, o = make_case(kind)
    ref = oracle.classify(p, s, mode=o["mode"], score_src=o["score_src"], skip_repeated=o["skip"], min_score=o["min_score"])
    assert ref["err"] == 0
    per_rank, outs = run_sharded(tmp_path, world, p, s, o)
    check_against_oracle(per_rank, outs, ref, world)
    if kind != "pe_cigar_blank_stop":
        assert not per_rank[0].get("declined")                  # the row kernels walked it: slivers, not a gather on rank 0


def test_sharded_walk_with_skewed_record_density(tmp_path):
    """the secondary stream's lines are much shorter in its second half: byte shards and record-index shards differ by
    hundreds of records, which travel as slivers (rows + text) between the ranks"""
    p, s, o = make_case("pe_liberal")
    lines = s.split(b"\n")[:-1]
    half = len(lines) // 2
    short = []
    for ln in lines[half:]:
        f = ln.split(b"\t")
        f[9] = f[9][:20]; f[10] = f[10][:20]                      # SEQ and QUAL cut to 20 bases
        short.append(b"\t".join(f))
    s2 = b"\n".join(lines[:half] + short) + b"\n"
    ref = oracle.classify(p, s2, mode=o["mode"], score_src=o["score_src"], skip_repeated=o["skip"], min_score=o["min_score"])
    per_rank, outs = run_sharded(tmp_path, 3, p, s2, o)
    check_against_oracle(per_rank, outs, ref, 3)
    assert max(r["sliver_bytes"] for r in per_rank) > 20000


def test_sharded_walk_reports_too_little_room(tmp_path):
    """slivers that do not fit the room around a shard end the walk on every rank with XM_ERR_NOMEM, not with garbage"""
    p, s, o = make_case("pe_liberal")
    lines = s.split(b"\n")[:-1]
    s2 = b"\n".join([b"\t".join(ln.split(b"\t")[:9] + [b"A", b"I"] + ln.split(b"\t")[11:]) for ln in lines[len(lines) // 2:]])
    s2 = b"\n".join(lines[:len(lines) // 2]) + b"\n" + s2 + b"\n"
    files = {}
    per_rank = None
    try:
        per_rank, _ = run_sharded(tmp_path, 2, p, s2, o, room=2048)
    except AssertionError:
        pytest.fail("a rank crashed instead of reporting the missing room")
    assert all(r["status"] == 6 for r in per_rank), per_rank


def test_single_rank_is_the_plain_walk(tmp_path):
    p, s, o = make_case("pe_liberal")
    ref = oracle.classify(p, s, mode=1)
    per_rank, outs = run_sharded(tmp_path, 1, p, s, o)
    check_against_oracle(per_rank, outs, ref, 1)


def test_byte_ranges_and_rendezvous(tmp_path):
    assert [sharded.byte_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert sharded.byte_range(0, 1, 3) == (0, 0)
    a, b = sharded.Rendezvous(0, 2, directory=str(tmp_path / "rv")), sharded.Rendezvous(1, 2, directory=str(tmp_path / "rv"))
    a.publish("blob", b"x" * 128)
    assert b.fetch("blob") == b"x" * 128
    with pytest.raises(TimeoutError):
        sharded.Rendezvous(1, 2, directory=str(tmp_path / "rv"), timeout=0.05).fetch("nothing")
    a.cleanup()
    assert not os.path.exists(str(tmp_path / "rv"))


def _gpu_count():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return n.value if cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["se_skip", "se_noskip_short_secondary", "pe_liberal", "pe_conservative_zs", "pe_cigar_blank_stop", "tiny"])
def test_sharded_entry_point_with_one_rank_on_gpu(tmp_path, kind):
    """xm_classify_sharded_host with a single rank: the whole record region is one (deliberately unaligned) shard --
    filler line, both scans into rows, k_size / k_prefix / k_emit, and the gather-to-rank-0 path for the blank line"""
    from xenomapper_b200 import _lib
    p, s, o = make_case(kind)
    ref = oracle.classify(p, s, mode=o["mode"], score_src=o["score_src"], skip_repeated=o["skip"], min_score=o["min_score"])
    ctx = _lib.Context(0)
    ctx.comm_init_rank(1, 0, None)
    opts = ctx.opts(o["mode"], o["score_src"], o["skip"], o["min_score"])
    rc, res, st, outs = ctx.classify_sharded_host(p, s, opts)
    assert rc == 0, ctx.error()
    assert list(res.counts) == ref["counts"] and int(res.n_records) == ref["n_yielded"]
    assert outs == ref["outputs"]
    assert list(st.out_total) == [len(x) for x in ref["outputs"]] and list(st.out_offset) == [0] * 6
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["se_skip", "pe_liberal", "pe_conservative_zs", "pe_cigar_blank_stop"])
def test_sharded_walk_on_two_gpus(tmp_path, kind):
    """two ranks, one device each, NCCL inside the library (`gpurun --gpus 2`)"""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    p, s, o = make_case(kind)
    ref = oracle.classify(p, s, mode=o["mode"], score_src=o["score_src"], skip_repeated=o["skip"], min_score=o["min_score"])
    per_rank, outs = run_sharded(tmp_path, 2, p, s, o, engine="gpu")
    check_against_oracle(per_rank, outs, ref, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [[], ["--paired"], ["--paired", "--conservative", "--min_score", "-18", "--use_zs"]],
                         ids=["se", "pe", "pe_conservative_zs"])
def test_sharded_cli_equals_single_process_cli(tmp_path, flags):
    """the launcher + `-m xenomapper_b200.xenomapper` (two ranks, one device each, NCCL inside the library) writes
    the same six files, headers included, as the single-process command"""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    style = synth.STYLE_PE_HISAT if "--use_zs" in flags else (synth.STYLE_PE_BOWTIE2 if flags else synth.STYLE_SE_BOWTIE2)
    p, s = synth.generate(3000, seed=21, style=style)
    if not flags:
        p, s = with_repeats(p), with_repeats(s, every=5)
    open(tmp_path / "p.sam", "wb").write(synth.HEADER_PRIMARY.encode() + bytes(p))
    open(tmp_path / "s.sam", "wb").write(synth.HEADER_SECONDARY.encode() + bytes(s))
    names = ["primary_specific", "secondary_specific", "primary_multi", "secondary_multi", "unassigned", "unresolved"]

    def run(tag, launcher, env):
        outs = []
        for n in names:
            outs += ["--" + n, str(tmp_path / ("%s_%s.sam" % (tag, n)))]
        cmd = launcher + ["-m", "xenomapper_b200.xenomapper", "--primary_sam", str(tmp_path / "p.sam"),
                          "--secondary_sam", str(tmp_path / "s.sam")] + outs + flags
        r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, **env), capture_output=True, timeout=600)
        assert r.returncode == 0, r.stderr.decode()[-3000:]
        return [open(tmp_path / ("%s_%s.sam" % (tag, n)), "rb").read() for n in names], r.stderr.decode()

    one, summary1 = run("one", [sys.executable], {})
    two, summary2 = run("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                "--master-addr", "127.0.0.1", "--master-port", str(free_port())],
                        {})
    assert two == one
    assert summary1[summary1.index("Read Count"):].strip() in summary2


@pytest.mark.gpu
def test_sharded_cli_over_nccl_on_two_gpus(tmp_path):
    """one process per GPU, NCCL for the words exchanged (needs two devices: `gpurun --gpus 2`)"""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    p, s = synth.generate(200000, seed=23, style=synth.STYLE_PE_BOWTIE2)
    open(tmp_path / "p.sam", "wb").write(synth.HEADER_PRIMARY.encode() + bytes(p))
    open(tmp_path / "s.sam", "wb").write(synth.HEADER_SECONDARY.encode() + bytes(s))
    names = ["primary_specific", "secondary_specific", "primary_multi", "secondary_multi", "unassigned", "unresolved"]
    outs = []
    for n in names:
        outs += ["--" + n, str(tmp_path / (n + ".sam"))]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), "-m", "xenomapper_b200.xenomapper", "--primary_sam", str(tmp_path / "p.sam"),
           "--secondary_sam", str(tmp_path / "s.sam"), "--paired"] + outs
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, timeout=900)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
    ref = oracle.classify(p, s, mode=1)
    for b, n in enumerate(names):
        data = open(tmp_path / (n + ".sam"), "rb").read()
        body = data[data.index(b"\n@CO") + 1:]
        body = body[body.index(b"\n") + 1:]
        assert body == ref["outputs"][b], n


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [[], ["--paired", "--cigar_scores", "--min_score", "-40"]], ids=["se", "pe_cigar"])
def test_sharded_cli_on_bam_inputs_equals_single_process_cli(tmp_path, flags):
    """--primary_bam / --secondary_bam under the launcher (two ranks, one device each): every rank maps both files, inflates
    and renders its part on its GPU (xm_bam_shard_*), the text shards go through the walk across GPUs.  The six files,
    headers included, equal those of the single-process command on the same BAM files."""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    from tests import _bamwriter
    from tests.test_bam import FULL_HEADER
    style = synth.STYLE_PE_BOWTIE2 if flags else synth.STYLE_SE_BOWTIE2
    p, s = synth.generate(60000, seed=29, style=style)
    if not flags:
        p, s = with_repeats(p), with_repeats(s, every=5)
    hdr2 = FULL_HEADER.replace("SN:chr", "SN:").replace("SN:M\t", "SN:MT\t")
    open(tmp_path / "p.bam", "wb").write(_bamwriter.sam_to_bam(FULL_HEADER, bytes(p)))
    open(tmp_path / "s.bam", "wb").write(_bamwriter.sam_to_bam(hdr2, bytes(s)))
    names = ["primary_specific", "secondary_specific", "primary_multi", "secondary_multi", "unassigned", "unresolved"]

    def run(tag, launcher):
        outs = []
        for n in names:
            outs += ["--" + n, str(tmp_path / ("%s_%s.sam" % (tag, n)))]
        cmd = launcher + ["-m", "xenomapper_b200.xenomapper", "--primary_bam", str(tmp_path / "p.bam"),
                          "--secondary_bam", str(tmp_path / "s.bam")] + outs + flags
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, timeout=600)
        assert r.returncode == 0, r.stderr.decode()[-3000:]
        return [open(tmp_path / ("%s_%s.sam" % (tag, n)), "rb").read() for n in names], r.stderr.decode()

    one, summary1 = run("one", [sys.executable])
    two, summary2 = run("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                "--master-addr", "127.0.0.1", "--master-port", str(free_port())])
    assert two == one
    assert summary1[summary1.index("Read Count"):].strip() in summary2
    assert sum(len(x) for x in one) > len(bytes(p)) // 2


# ---- the ranks' BAM record chains joined (sharded._settle_entries), without a GPU ---------------------------------------------
class _FakeChainCtx:
    """what xm_bam_shard_open / xm_bam_shard_chain do, on a list of record offsets: a rank's part is [lo, hi) of the inflated
    stream; its exit is the first record start at or behind hi; chained from an entry at or behind hi it answers with the entry"""
    NONE64 = (1 << 64) - 1

    def __init__(self, starts, total, lo, hi, guess):
        self.starts, self.total, self.lo, self.hi, self.guess = starts, total, lo, hi, guess
        self.calls = 0

    def _exit(self):
        nxt = [s for s in self.starts if s >= self.hi]
        return nxt[0] if nxt else self.total

    def open(self):
        return (self.guess, self._exit()) if self.guess != self.NONE64 else (self.NONE64, self.NONE64)

    def bam_shard_chain(self, stream, entry):
        self.calls += 1
        self.entry = entry
        return entry if entry >= self.hi else self._exit()


@pytest.mark.parametrize("world", [2, 3, 5, 8])
@pytest.mark.parametrize("scenario", ["all right", "one wrong", "all wrong", "nothing seen", "header parts", "long record"])
def test_bam_chain_settles_between_ranks(tmp_path, world, scenario):
    """every rank follows its part from the exit of the ranks before it, whatever it guessed at first"""
    import random
    import threading
    from xenomapper_b200 import sharded
    rnd = random.Random(world * 7 + len(scenario))
    total = 100000
    first_record = 20000 if scenario == "header parts" else 137
    starts, p = [], first_record
    while p < total:
        starts.append(p)
        p += 60000 if (scenario == "long record" and len(starts) == 3) else rnd.randrange(40, 900)
    bounds = [total * r // world for r in range(world + 1)]
    true_entry = []
    for r in range(world):
        nxt = [s for s in starts if s >= bounds[r]]
        true_entry.append(nxt[0] if nxt else total)
    ctxs = []
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        own = [s for s in starts if lo <= s < hi]
        if not own:
            guess = _FakeChainCtx.NONE64                                         # nothing that looks like a record start
        elif r == 0 or scenario == "all right" or (scenario in ("header parts", "long record", "nothing seen")):
            guess = own[0]
        elif scenario == "one wrong":
            guess = own[0] if r != world // 2 else own[min(1, len(own) - 1)] + 1
        else:
            guess = own[min(1, len(own) - 1)] + 3                                # a false positive further in
        if scenario == "nothing seen" and r == world - 1 and own:
            guess = _FakeChainCtx.NONE64                                         # the guesser found nothing although records start there
        if r == 0 and own:
            guess = own[0]                                                       # rank 0 knows where the records start
        ctxs.append(_FakeChainCtx(starts, total, lo, hi, guess))
    errors = []

    def run(r):
        try:
            rv = sharded.Rendezvous(r, world, directory=str(tmp_path / "rv"), timeout=30)
            g, e = ctxs[r].open()
            sharded._settle_entries(ctxs[r], r, world, rv, 0, g, e, "t")
        except Exception as ex:                                                  # noqa: BLE001
            errors.append((r, repr(ex)))

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join(60) for t in th]
    assert not errors, errors
    for r in range(world):
        c = ctxs[r]
        used = getattr(c, "entry", c.guess)
        own = [s for s in starts if bounds[r] <= s < bounds[r + 1]]
        if used == c.NONE64:
            # a part that passes the chain on: fine only while no rank before it holds records (header) -- or it holds none itself
            assert not own or all(not [s for s in starts if bounds[q] <= s < bounds[q + 1]] for q in range(r)) and not own
        elif own:
            assert used == true_entry[r], (r, used, true_entry[r])
        else:
            assert used >= bounds[r + 1] or used == total                        # the chain jumps over the part
