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


def run_sharded(tmp_path, world, p, s, o, debug=0, engine="emu"):
    files = {}
    for name, data in (("prim", p), ("sec", s)):
        files[name] = str(tmp_path / (name + ".sam"))
        open(files[name], "wb").write(data)
    outs = [str(tmp_path / ("out%d.sam" % b)) for b in range(6)]
    for f in outs:
        open(f, "wb").close()
    case = dict(files, outs=outs, result=str(tmp_path / "result.json"), enabled_bins=0x3F, debug=debug, engine=engine, **o)
    casefile = str(tmp_path / "case.json")
    json.dump(case, open(casefile, "w"))
    port = free_port()
    procs = [subprocess.Popen([sys.executable, "-m", "tests._sharded_worker", str(r), str(world), str(port), casefile],
                              cwd=ROOT, stderr=subprocess.PIPE) for r in range(world)]
    for pr in procs:
        _, err = pr.communicate(timeout=600)
        assert pr.returncode == 0, err.decode()[-2000:]
    return json.load(open(case["result"])), [open(f, "rb").read() for f in outs]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("kind", ["se_skip", "se_noskip_short_secondary", "pe_liberal", "pe_conservative_zs",
                                  "pe_cigar_blank_stop", "tiny"])
def test_sharded_walk_equals_single_pass(tmp_path, kind, world):
    p, s, o = make_case(kind)
    ref = oracle.classify(p, s, mode=o["mode"], score_src=o["score_src"], skip_repeated=o["skip"], min_score=o["min_score"])
    assert ref["err"] == 0
    res, outs = run_sharded(tmp_path, world, p, s, o)
    assert res["status"] == 0, res["message"]
    assert res["counts"] == ref["counts"]
    assert res["n_records"] == ref["n_yielded"]
    assert res["out_total"] == [len(x) for x in ref["outputs"]]
    assert outs == ref["outputs"]


def test_sharded_walk_small_tiles_many_boundaries(tmp_path):
    """1 KiB tiles inside every shard on top of the shard boundaries"""
    p, s, o = make_case("se_skip")
    p, s = p[:60000], s[:50000]
    p, s = p[:p.rfind(b"\n") + 1], s[:s.rfind(b"\n") + 1]
    ref = oracle.classify(p, s, mode=0, skip_repeated=True)
    res, outs = run_sharded(tmp_path, 2, p, s, o, debug=2)
    assert res["status"] == 0, res["message"]
    assert res["counts"] == ref["counts"] and outs == ref["outputs"]


def test_single_rank_is_the_plain_walk(tmp_path):
    p, s, o = make_case("pe_liberal")
    ref = oracle.classify(p, s, mode=1)
    res, outs = run_sharded(tmp_path, 1, p, s, o)
    assert res["counts"] == ref["counts"] and outs == ref["outputs"]


def test_line_alignment_helpers():
    src = sharded.BytesSource(b"aa\nbbbb\n\ncc\n")
    assert [sharded.line_start_at_or_after(src, x) for x in range(13)] == [0, 3, 3, 3, 8, 8, 8, 8, 8, 9, 12, 12, 12]
    assert sharded.previous_line_start(src, 3) == 0
    assert sharded.previous_line_start(src, 8) == 3
    assert sharded.previous_line_start(src, 9) == 8
    assert sharded.plan_partition(10, 4) == [0, 2, 5, 7, 10]


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["se_skip", "pe_liberal", "pe_conservative_zs", "pe_cigar_blank_stop"])
def test_sharded_walk_on_gpu(tmp_path, kind):
    """two ranks (gloo for the few words exchanged), each driving the CUDA kernels through the C ABI"""
    p, s, o = make_case(kind)
    ref = oracle.classify(p, s, mode=o["mode"], score_src=o["score_src"], skip_repeated=o["skip"], min_score=o["min_score"])
    res, outs = run_sharded(tmp_path, 2, p, s, o, engine="gpu")
    assert res["status"] == 0, res["message"]
    assert res["counts"] == ref["counts"]
    assert outs == ref["outputs"]


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [[], ["--paired"], ["--paired", "--conservative", "--min_score", "-18", "--use_zs"]],
                         ids=["se", "pe", "pe_conservative_zs"])
def test_sharded_cli_equals_single_process_cli(tmp_path, flags):
    """`torchrun -m xenomapper_b200.xenomapper` (two ranks on one device, gloo for the words exchanged) writes the
    same six files, headers included, as the single-process command"""
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
                        {"XENOMAPPER_DIST_BACKEND": "gloo", "XENOMAPPER_DEVICE": "0"})
    assert two == one
    assert summary1[summary1.index("Read Count"):].strip() in summary2


@pytest.mark.gpu
def test_sharded_cli_over_nccl_on_two_gpus(tmp_path):
    """one process per GPU, NCCL for the words exchanged (needs two devices: `gpurun --gpus 2`)"""
    import torch
    if torch.cuda.device_count() < 2:
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
