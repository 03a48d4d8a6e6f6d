#!/usr/bin/env python3
# encoding: utf-8
"""
xenomapper_b200.xenomapper -- host side of the B200 read-binning path.

Same names, arguments, return values and exceptions as
xenomapper/xenomapper.py of genomematt/xenomapper v1.0.2 ("xm.py"), so a
caller (or the reference's own test-suite) can switch modules without other
changes.  What differs is where the work happens: the lockstep walk of
getReadPairs + main_single_end / main_paired_end /
conservative_main_paired_end (xm.py:95-118, 291-556) is ONE call into
libxenomapper_b200.so (include/xenomapper_b200.h), which runs the sm_100a CUDA
kernels.  There is no Python or CPU implementation of the walk in this module:
without the library and a B200 the walk raises.

What stays in Python, because it is text glue around the walk and not part of
it: SAM header handling (xm.py:36-46, 120-174), the summary table
(xm.py:558-566) and the command line (xm.py:568-743).  get_tag*,
get_cigarbased_AS_tag and get_mapping_state are kept as single-record
utilities and as the `tag_func` selectors of the walks; the walks never call
them (tags are parsed in-kernel).
"""
import argparse
import io
import os
import re
import sys
import stat
import textwrap
import weakref
from collections import Counter

from . import _lib
from ._lib import UnsupportedInput, XenomapperLibraryError  # noqa: F401  (re-exported)

__version__ = "1.0.2"        # the reference version this module mirrors; written into @PG lines (xm.py:128)

STATES = _lib.BINS
NEG_INF = float("-inf")


# ---------------------------------------------------------------------------
# headers (boundary glue; byte parity "headers included")

def get_sam_header(samfile):
    """Leading '@' lines of a SAM file, without newlines; leaves the file at the
    first record.  Mirrors xm.py:36-46, including IndexError on a header-only
    or empty file."""
    header = []
    while True:
        mark = samfile.tell()
        line = samfile.readline().strip('\n')
        if line[0] != '@':          # '' raises IndexError, as in the reference
            samfile.seek(mark)
            return header
        header.append(line)


def get_bam_header(bamfile):
    """Header text lines stored in a BAM file (the `samtools view -H` of xm.py:48-54)."""
    from . import bam
    text = bam.read_header_text(bamfile)
    bamfile.seek(0)
    return [x for x in text.split('\n') if x]


def add_pg_tag(sam_header_list, comment=None):
    """Header plus Xenomapper's @PG line (chained with PP: to a trailing @PG) and an
    optional @CO line.  Mirrors xm.py:120-131."""
    out = list(sam_header_list)
    if any(not h.startswith('@') for h in out):
        raise ValueError('Incorrect SAM header format :\n{0}'.format('\n'.join(out)))
    pp = ''
    if out[-1][:3] == '@PG':        # IndexError on an empty header, as in the reference
        ids = [t for t in out[-1].split() if t[:2] == 'ID']
        pp = 'PP' + ids[0][2:] + '\t'
    out.append('@PG\tID:Xenomapper\tPN:Xenomapper\t' + pp + 'VN:{0}'.format(__version__))
    if comment:
        out.append('@CO\t' + comment)
    return out


_RECORD_OFFSET = weakref.WeakKeyDictionary()     # input file -> byte offset of its first record, as the library found it


def _regular_fd(f):
    try:
        fd = f.fileno()
        return fd if stat.S_ISREG(os.fstat(fd).st_mode) else None
    except (AttributeError, OSError, ValueError):
        return None


def _process_headers_native(file1, file2, outs):
    """process_headers on the raw bytes of two regular files, inside the library (xm_process_headers_fds): headers,
    @PG / @CO lines and the byte offset of the first record, without tell() cookies.  Returns False when the inputs
    are not plain files at their start (the Python text path below handles those)."""
    fds = [_regular_fd(file1), _regular_fd(file2)]
    if None in fds:
        return False
    try:
        if file1.tell() != 0 or file2.tell() != 0:
            return False
    except (OSError, ValueError):
        return False
    rc, bad, offsets, texts, status = _lib.process_headers(fds[0], fds[1], __version__)
    if rc == _lib.XM_ERR_INDEX:
        raise IndexError('string index out of range')                    # ''[0] in get_sam_header, xm.py:40-43
    if rc == _lib.XM_ERR_UNICODE:
        raise UnicodeDecodeError('utf-8', b'', 0, 1, 'invalid start byte in the header of input {0}'.format(bad + 1))
    if rc != _lib.XM_OK:
        raise XenomapperLibraryError('header processing failed ({0})'.format(rc))
    for k, f in enumerate((file1, file2)):
        f.seek(offsets[k])                                               # the reference leaves both files at their first record
        _RECORD_OFFSET[f] = offsets[k]
    for b, out in enumerate(outs):
        if b and not out:
            continue                                                     # primary_specific is always printed (None = stdout)
        if status[b] == _lib.XM_ERR_INDEX:
            raise IndexError('list index out of range')                 # add_pg_tag on an empty header / @PG without ID, xm.py:124-127
        print(texts[b].decode('utf-8'), end='', file=out)
    return True


def process_headers(file1, file2, primary_specific=sys.stdout, secondary_specific=None,
                    primary_multi=None, secondary_multi=None, unassigned=None, unresolved=None, bam=False):
    """Write each enabled output's header.  Mirrors xm.py:133-174: primary-species
    header for primary_*, unassigned and unresolved; secondary-species header
    for secondary_*.  Plain SAM files at their start go through the library."""
    outs = (primary_specific, secondary_specific, primary_multi, secondary_multi, unassigned, unresolved)
    if not bam and _process_headers_native(file1, file2, outs):
        return
    reader = get_bam_header if bam else get_sam_header
    h1, h2 = reader(file1), reader(file2)
    print('\n'.join(add_pg_tag(h1, comment='species specific reads')), file=primary_specific)
    plan = ((secondary_specific, h2, 'species specific reads'),
            (primary_multi, h1, 'species specific multimapping reads'),
            (secondary_multi, h2, 'species specific multimapping reads'),
            (unassigned, h1, 'reads that could not be assigned'),
            (unresolved, h1, 'reads that could not be resolved'))
    for out, hdr, comment in plan:
        if out:
            print('\n'.join(add_pg_tag(hdr, comment=comment)), file=out)


# ---------------------------------------------------------------------------
# single-record utilities and tag_func selectors (not used by the walks)

def get_tag(sam_line, tag='AS'):
    """Value of a numeric SAM tag, -inf if absent.  Mirrors xm.py:176-191
    (substring match on fields >= 11, ValueError on several matches)."""
    hits = [f for f in sam_line[11:] if tag in f]
    if not hits:
        return NEG_INF
    if len(hits) > 1:
        raise ValueError('SAM line has multiple values of {0}: {1}'.format(tag, sam_line))
    return float(hits[0].split(':')[-1])


def get_tag_with_ZS_as_XS(sam_line, tag='AS'):
    """get_tag, answering requests for XS with the ZS tag (HISAT).  Mirrors xm.py:193-206."""
    return get_tag(sam_line, 'ZS' if tag == 'XS' else tag)


_CIGAR_OP = re.compile(r'([0-9]+)([MIDNSHPX=])')
_UNIVERSAL_NEWLINE = re.compile(r'\r\n|\r|\n')


def get_cigarbased_AS_tag(sam_line, tag='AS'):
    """Score derived from CIGAR and NM: -6/mismatch, -5/gap open, -3/gap base,
    -2/soft-clipped base.  Other tags come from get_tag.  Mirrors xm.py:228-256."""
    if tag != 'AS':
        return get_tag(sam_line, tag)
    nm = [f for f in sam_line[11:] if 'NM' in f]
    if not nm:
        return NEG_INF
    mismatches = int(nm[0].split(':')[-1])
    gaps, gap_bases, clipped = 0, 0, 0
    for n, op in _CIGAR_OP.findall(sam_line[5]):
        if op in 'ID':
            gaps += 1
            gap_bases += int(n)
        elif op == 'S':
            clipped += int(n)
    return -6 * mismatches - 5 * gaps - 3 * gap_bases - 2 * clipped


def get_mapping_state(AS1, XS1, AS2, XS2, min_score=NEG_INF):
    """The six-way decision of xm.py:258-289 for one read."""
    if AS1 <= min_score and AS2 <= min_score:
        return 'unassigned'
    if AS1 > min_score and (AS2 <= min_score or AS1 > AS2):
        return 'primary_specific' if (not XS1 or AS1 > XS1) else 'primary_multi'
    if AS1 == AS2:
        return 'unresolved'
    if AS2 > min_score and (AS1 <= min_score or AS2 > AS1):
        return 'secondary_specific' if (not XS2 or AS2 > XS2) else 'secondary_multi'
    raise RuntimeError('Error in processing logic with values {0} '.format((AS1, XS1, AS2, XS2)))


_SCORE_SRC = {get_tag: _lib.SCORE_AS_XS, get_tag_with_ZS_as_XS: _lib.SCORE_AS_ZS,
              get_cigarbased_AS_tag: _lib.SCORE_CIGAR_NM}


# ---------------------------------------------------------------------------
# the lockstep reader: a handle on the two record regions

class ReadPairs:
    """What getReadPairs returns: the two inputs positioned after their headers,
    plus the skip flag.  main_single_end / main_paired_end /
    conservative_main_paired_end recognise it and hand both record regions to
    the library in one call.  Iterating it yields (fields1, fields2) like the
    reference's generator, for callers that consume the pairs themselves."""

    def __init__(self, sam1, sam2, skip_repeated_reads=False, bam=False):
        self.sam1, self.sam2 = sam1, sam2
        self.skip_repeated_reads = bool(skip_repeated_reads)
        self.bam = bam

    def record_regions(self):
        """bytes of both record regions, through the inputs' own text layer
        (so decoding errors and universal newlines behave as in the reference)"""
        if self.bam:
            from . import bam
            return bam.records_as_sam_text(self.sam1), bam.records_as_sam_text(self.sam2)
        return _remaining_bytes(self.sam1), _remaining_bytes(self.sam2)

    def __iter__(self):
        p, s = self.record_regions()
        # lines end as the reference's 'rt' files end them (universal newlines: "\n", "\r\n" or a lone "\r")
        l1 = iter(_UNIVERSAL_NEWLINE.split(p.decode('utf-8', 'surrogateescape')))
        l2 = iter(_UNIVERSAL_NEWLINE.split(s.decode('utf-8', 'surrogateescape')))
        a, b = next(l1, '').split(), next(l2, '').split()
        while a and b:
            assert a[0] == b[0]
            yield a, b
            if self.skip_repeated_reads:
                na, nb = a[0], b[0]
                while a and b and a[0] == na:
                    a = next(l1, '').split()
                while a and b and b[0] == nb:
                    b = next(l2, '').split()
            else:
                a, b = next(l1, '').split(), next(l2, '').split()


def _remaining_bytes(f):
    data = f.read()
    if isinstance(data, str):
        data = data.encode('utf-8', 'surrogateescape')
    return data


def getReadPairs(sam1, sam2, skip_repeated_reads=False):
    """Pair up the records of two SAM files read in lockstep (xm.py:95-118)."""
    return ReadPairs(sam1, sam2, skip_repeated_reads)


def getBamReadPairs(bamfile1, bamfile2, skip_repeated_reads=False):
    """Pair up the records of two BAM files (xm.py:66-93), decoded without samtools."""
    return ReadPairs(bamfile1, bamfile2, skip_repeated_reads, bam=True)


def _serialise_pairs(readpairs):
    """A caller-supplied iterable of (fields1, fields2): back to two SAM record regions."""
    p, s = [], []
    for a, b in readpairs:
        for fields in (a, b):
            if not fields or any((not t) or len(t.split()) != 1 or t != t.strip() for t in fields):
                raise UnsupportedInput("field lists with empty or whitespace-bearing fields cannot be re-serialised")
        p.append('\t'.join(a))
        s.append('\t'.join(b))
    enc = lambda rows: ('\n'.join(rows) + '\n').encode('utf-8', 'surrogateescape') if rows else b''
    return enc(p), enc(s)


# ---------------------------------------------------------------------------
# the walks

def _raise_for(rc, ctx, res):
    msg = ctx.error()
    if rc == _lib.XM_ERR_ASSERT:
        raise AssertionError(msg)
    if rc == _lib.XM_ERR_VALUE:
        raise ValueError(msg)
    if rc == _lib.XM_ERR_RUNTIME:
        raise RuntimeError(msg)
    if rc == _lib.XM_ERR_UNICODE:
        raise UnicodeDecodeError('utf-8', b'', 0, 1, msg)
    if rc == _lib.XM_ERR_UNSUPPORTED:
        raise UnsupportedInput(msg)
    raise XenomapperLibraryError(msg)


def _counter(res, paired):
    c = Counter()
    if paired:
        for f in range(6):
            for r in range(6):
                n = res.counts[f * 6 + r]
                if n:
                    c[(STATES[f], STATES[r])] = n
    else:
        for k in range(6):
            if res.counts[k]:
                c[STATES[k]] = res.counts[k]
    return c


def _walk(mode, readpairs, outputs, min_score, tag_func):
    try:
        score_src = _SCORE_SRC[tag_func]
    except (KeyError, TypeError):
        raise NotImplementedError(
            "tag_func must be get_tag, get_tag_with_ZS_as_XS or get_cigarbased_AS_tag: scores are parsed "
            "inside the CUDA kernels, arbitrary Python callables cannot run there")
    bam_inputs = fds = None
    if isinstance(readpairs, ReadPairs) and readpairs.bam:
        from . import bam
        bam_inputs = (bam._all_bytes(readpairs.sam1), bam._all_bytes(readpairs.sam2))
        skip = readpairs.skip_repeated_reads
    elif isinstance(readpairs, ReadPairs):
        skip = readpairs.skip_repeated_reads
        fds = _descriptors(readpairs, outputs)      # before anything is read from the inputs
        if not fds:
            if getattr(readpairs, "bgzf", False):
                raise NotImplementedError("BGZF output needs regular input files and outputs with descriptors")
            prim, sec = readpairs.record_regions()
    else:
        prim, sec = _serialise_pairs(readpairs)
        skip = False
    enabled = 0
    for b, out in enumerate(outputs):
        if out:                      # the reference tests truthiness (xm.py:333)
            enabled |= 1 << b
    ctx = _lib.default_context()
    opts = ctx.opts(mode, score_src, skip, float(min_score), enabled)
    if fds:
        # real files on both sides (the CLI): the library streams them through pinned staging and appends each bin to
        # its descriptor in record order (xm_classify_fds); nothing passes through Python
        (fd1, off1), (fd2, off2), out_fds = fds
        rc, res = ctx.classify_fds(fd1, off1, fd2, off2, out_fds, opts, _lib.OUT_BGZF if getattr(readpairs, "bgzf", False) else 0)
        if rc != _lib.XM_OK:
            _raise_for(rc, ctx, res)
        return _counter(res, mode != _lib.MODE_SE)
    if bam_inputs:
        out_fds = _output_descriptors(outputs)
        if out_fds:
            # BAM in, descriptors out: inflate, record chain, rendering, the walk (and the deflate of --bgzf) on the GPU
            rc, res = ctx.classify_bam_fds(bam_inputs[0], bam_inputs[1], out_fds, opts,
                                           _lib.OUT_BGZF if getattr(readpairs, "bgzf", False) else 0)
            if rc != _lib.XM_OK:
                _raise_for(rc, ctx, res)
            return _counter(res, mode != _lib.MODE_SE)
        if getattr(readpairs, "bgzf", False):
            raise NotImplementedError("BGZF output needs outputs with descriptors")
        rc, res, outs = ctx.classify_bam_host(bam_inputs[0], bam_inputs[1], opts)
    else:
        rc, res, outs = ctx.classify_host(prim, sec, opts)
    # everything before a failing record is written, like the reference's streaming prints
    for b, out in enumerate(outputs):
        if out and outs[b]:
            out.write(outs[b].decode('utf-8', 'surrogateescape'))
    if rc != _lib.XM_OK:
        _raise_for(rc, ctx, res)
    return _counter(res, mode != _lib.MODE_SE)


def _descriptors(readpairs, outputs):
    """((fd, byte offset) of both inputs, six output descriptors) when every file involved is a real one, else None.
    Text inputs must sit where only whole lines were read (after process_headers): then tell() is the byte offset
    of the first record, which is checked against the raw bytes."""
    ins = []
    for f in (readpairs.sam1, readpairs.sam2):
        try:
            fd = f.fileno()
            if not stat.S_ISREG(os.fstat(fd).st_mode):
                return None
            pos = f.tell()
        except (AttributeError, OSError, ValueError):
            return None
        if _RECORD_OFFSET.get(f) == pos:
            ins.append((fd, pos))             # where the library's header pass left it: a byte offset by construction
            continue
        if not isinstance(pos, int) or pos < 0 or pos > os.fstat(fd).st_size:
            return None
        if pos > 0 and os.pread(fd, 1, pos - 1) != b"\n":
            return None                      # not at a line start: a decoder-state cookie, not an offset
        if getattr(f, "newlines", None) not in (None, "\n"):
            return None                      # universal newlines already translated something: let the text layer decide
        ins.append((fd, pos))
    out_fds = _output_descriptors(outputs)
    if out_fds is None:
        return None
    return ins[0], ins[1], out_fds


def _output_descriptors(outputs):
    """the six outputs' descriptors (-1: disabled), flushed; None when one of them has none"""
    out_fds = []
    for out in outputs:
        if not out:
            out_fds.append(-1)
            continue
        try:
            out.flush()                      # header lines written by process_headers
            out_fds.append(out.fileno())
        except (AttributeError, OSError, ValueError):
            return None
    return out_fds


def main_single_end(readpairs, primary_specific=sys.stdout, secondary_specific=None, primary_multi=None,
                    secondary_multi=None, unassigned=None, unresolved=None, min_score=NEG_INF, tag_func=get_tag):
    """Bin single-end reads (xm.py:291-352).  Returns a Counter keyed by category."""
    outs = (primary_specific, secondary_specific, primary_multi, secondary_multi, unassigned, unresolved)
    return _walk(_lib.MODE_SE, readpairs, outs, min_score, tag_func)


def main_paired_end(readpairs, primary_specific=sys.stdout, secondary_specific=None, primary_multi=None,
                    secondary_multi=None, unassigned=None, unresolved=None, min_score=NEG_INF, tag_func=get_tag):
    """Bin interlaced read pairs, liberal priority chain (xm.py:354-454).
    Returns a Counter keyed by (forward, reverse) category."""
    outs = (primary_specific, secondary_specific, primary_multi, secondary_multi, unassigned, unresolved)
    return _walk(_lib.MODE_PE_LIBERAL, readpairs, outs, min_score, tag_func)


def conservative_main_paired_end(readpairs, primary_specific=sys.stdout, secondary_specific=None, primary_multi=None,
                                 secondary_multi=None, unassigned=None, unresolved=None, min_score=NEG_INF,
                                 tag_func=get_tag):
    """Bin interlaced read pairs, conservative chain (xm.py:456-556)."""
    outs = (primary_specific, secondary_specific, primary_multi, secondary_multi, unassigned, unresolved)
    return _walk(_lib.MODE_PE_CONSERVATIVE, readpairs, outs, min_score, tag_func)


def output_summary(category_counts, outfile=sys.stderr):
    """The markdown table of xm.py:558-566."""
    w = lambda *a, **k: print(*a, file=outfile, **k)
    w('-' * 80)
    w('Read Count Category Summary\n')
    w('|       {0:45s}|     {1:10s}  |'.format('Category', 'Count'))
    w('|:' + '-' * 50 + ':|:' + '-' * 15 + ':|')
    for category in sorted(category_counts):
        w('|  {0:50s}|{1:15d}  |'.format(str(category), category_counts[category]))
    w()


# ---------------------------------------------------------------------------
# command line (same flags as xm.py:568-678)

def command_line_interface(argv=None):
    p = argparse.ArgumentParser(
        prog="xenomapper", formatter_class=argparse.RawDescriptionHelpFormatter,
        description=textwrap.dedent("""\
            Sorts the reads of two SAM (or BAM) files, the same reads aligned to a primary and a
            secondary species, into species specific, multimapping, unresolved and unassigned
            outputs.  B200 build: the read-binning walk runs on the GPU.

            Both files need AS and XS scores where higher is better (Bowtie2 --local; with -p also
            --reorder), the reads in the same order.  --cigar_scores and --use_zs cover aligners
            without usable AS/XS tags.  Inputs must be seekable."""),
        epilog="To write BAM use process substitution:  --primary_specific >(samtools view -bS - > out.bam)")
    a = p.add_argument
    a('--primary_sam', type=argparse.FileType('rt'), default=None, help='SAM file of the alignment to the primary species')
    a('--secondary_sam', type=argparse.FileType('rt'), default=None, help='SAM file of the alignment to the secondary species')
    a('--primary_bam', type=argparse.FileType('rb'), default=None, help='BAM file of the alignment to the primary species')
    a('--secondary_bam', type=argparse.FileType('rb'), default=None, help='BAM file of the alignment to the secondary species')
    a('--primary_specific', type=argparse.FileType('wt'), default=sys.stdout, help='output: reads specific to the primary species')
    a('--secondary_specific', type=argparse.FileType('wt'), default=None, help='output: reads specific to the secondary species')
    a('--primary_multi', type=argparse.FileType('wt'), default=None, help='output: reads multimapping in the primary species')
    a('--secondary_multi', type=argparse.FileType('wt'), default=None, help='output: reads multimapping in the secondary species')
    a('--unassigned', type=argparse.FileType('wt'), default=None, help='output: reads mapping in neither species')
    a('--unresolved', type=argparse.FileType('wt'), default=None, help='output: reads mapping equally well in both species')
    a('--paired', action='store_true', help='reads are paired, mates interlaced, each pair once')
    a('--conservative', action='store_true', help='paired reads: any unassigned mate makes the pair unassigned, discordant species unresolved')
    a('--min_score', type=float, default=NEG_INF, help='scores less than or equal to this count as unmapped')
    a('--cigar_scores', action='store_true', help='score = -6*NM -5*gap opens -3*gap bases -2*soft clipped bases, from CIGAR and NM')
    a('--use_zs', action='store_true', help='take the next-best score from ZS instead of XS (HISAT)')
    a('--bgzf', action='store_true', help='write the outputs as BGZF (blocked gzip of the SAM text, what bgzip writes) instead of '
                                          'plain SAM; not an option of the reference, which leaves compression to a pipe')
    a('--version', action='store_true', help='print version information and exit')
    args = p.parse_args(argv)
    if args.version:
        print(__version__)
        sys.exit()
    if (not args.primary_sam or not args.secondary_sam) and (not args.primary_bam or not args.secondary_bam):
        print('ERROR: You must provide --primary_sam and --secondary_sam\n or --primary_bam and --secondary_bam\n')
        p.print_help()
        sys.exit(1)
    return args


def main(argv=None):
    """Console entry point (xm.py:681-743)."""
    args = command_line_interface(argv)
    tag_func = get_cigarbased_AS_tag if args.cigar_scores else (get_tag_with_ZS_as_XS if args.use_zs else get_tag)
    outs = dict(primary_specific=args.primary_specific, secondary_specific=args.secondary_specific,
                primary_multi=args.primary_multi, secondary_multi=args.secondary_multi,
                unassigned=args.unassigned, unresolved=args.unresolved)
    skip = not args.paired                       # xm.py:691
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return _main_sharded(args, tag_func, outs, skip)
    if args.primary_sam and (_regular_fd(args.primary_sam) is None or _regular_fd(args.secondary_sam) is None):
        return _main_streams(args, tag_func, outs, skip)
    if args.bgzf:
        return _main_bgzf(args, tag_func, outs, skip)
    if args.primary_sam:
        process_headers(args.primary_sam, args.secondary_sam, **outs)
        pairs = getReadPairs(args.primary_sam, args.secondary_sam, skip_repeated_reads=skip)
    else:
        process_headers(args.primary_bam, args.secondary_bam, bam=True, **outs)
        pairs = getBamReadPairs(args.primary_bam, args.secondary_bam, skip_repeated_reads=skip)
    walk = main_single_end if not args.paired else (conservative_main_paired_end if args.conservative else main_paired_end)
    counts = walk(pairs, min_score=args.min_score, tag_func=tag_func, **outs)
    for f in outs.values():
        if f:
            f.flush()
    output_summary(counts)


def _main_streams(args, tag_func, outs, skip):
    """Inputs that cannot seek -- pipes from two aligners, process substitutions (the reference refuses them,
    xm.py:586-587): headers and walk in one library call on the descriptors themselves (xm_classify_streams)."""
    files = [outs[k] for k in outs]
    fds = []
    for f in files:
        if f:
            f.flush()
            fds.append(f.fileno())
        else:
            fds.append(-1)
    mode = _lib.MODE_SE if not args.paired else (_lib.MODE_PE_CONSERVATIVE if args.conservative else _lib.MODE_PE_LIBERAL)
    ctx = _lib.default_context()
    opts = ctx.opts(mode, _SCORE_SRC[tag_func], skip, float(args.min_score), sum(1 << b for b, f in enumerate(fds) if f >= 0))
    rc, res = ctx.classify_streams(args.primary_sam.fileno(), args.secondary_sam.fileno(), fds, opts, __version__,
                                   _lib.OUT_BGZF if args.bgzf else 0)
    if rc == _lib.XM_ERR_INDEX:
        raise IndexError('string index out of range')
    if rc != _lib.XM_OK:
        _raise_for(rc, ctx, res)
    output_summary(_counter(res, mode != _lib.MODE_SE))


def _main_bgzf(args, tag_func, outs, skip):
    """--bgzf: headers and bins leave as BGZF members (xm_bgzf_write, xm_classify_fds_ex with XM_OUT_BGZF).
    gunzip of every output equals what the command writes without the flag."""
    import io
    bam = not args.primary_sam
    files = {k: f for k, f in outs.items() if f}
    for k, f in files.items():
        if _regular_fd(f) is None and not hasattr(f, "fileno"):
            raise ValueError("--bgzf writes through descriptors: {0} has none".format(k))
    texts = {k: io.StringIO() for k in files}
    if bam:
        process_headers(args.primary_bam, args.secondary_bam, bam=True, **{k: texts.get(k) for k in outs})
    else:
        process_headers(args.primary_sam, args.secondary_sam, **{k: texts.get(k) for k in outs})
    for k, f in files.items():
        f.flush()
        _lib.bgzf_write(f.fileno(), texts[k].getvalue().encode())
    if bam:
        pairs = getBamReadPairs(args.primary_bam, args.secondary_bam, skip_repeated_reads=skip)
    else:
        pairs = getReadPairs(args.primary_sam, args.secondary_sam, skip_repeated_reads=skip)
    pairs.bgzf = True
    walk = main_single_end if not args.paired else (conservative_main_paired_end if args.conservative else main_paired_end)
    try:
        counts = walk(pairs, min_score=args.min_score, tag_func=tag_func, **outs)
    finally:
        for f in files.values():
            _lib.bgzf_write(f.fileno(), eof=True)
    output_summary(counts)


def _main_sharded(args, tag_func, outs, skip):
    """Started once per GPU by a launcher that sets RANK, WORLD_SIZE and LOCAL_RANK: each process walks the
    records of its byte shard (xenomapper_b200/sharded.py, csrc/xm_shard.h; NCCL inside the library).  Outputs must be regular files: every rank writes
    its part of each bin in place behind the header."""
    import json
    from . import sharded
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    names = list(outs)
    files = [outs[k] for k in names]
    for f in files:
        if f and not os.path.isfile(f.name):
            raise ValueError("sharded walks write their outputs in place: {0} is not a regular file".format(f.name))
    ctx = _lib.Context(local)
    rv = sharded.init_comm(ctx, rank, world)
    ctx.comm_barrier()                              # every rank has opened (and truncated) the outputs
    bam = not args.primary_sam
    hdr_len = [0] * 6
    if rank == 0:
        if bam:
            process_headers(args.primary_bam, args.secondary_bam, bam=True, **outs)
        else:
            process_headers(args.primary_sam, args.secondary_sam, **outs)
        for b, f in enumerate(files):
            if f:
                f.flush()
                hdr_len[b] = os.fstat(f.fileno()).st_size
        rv.publish("header_len", json.dumps(hdr_len).encode())
    else:
        hdr_len = json.loads(rv.fetch("header_len").decode())
    mode = _lib.MODE_SE if not args.paired else (_lib.MODE_PE_CONSERVATIVE if args.conservative else _lib.MODE_PE_LIBERAL)
    enabled = sum(1 << b for b, f in enumerate(files) if f)
    if bam:
        # every rank maps both files; its part of each is inflated and rendered on its GPU (xm_bam_shard_*)
        from . import bam as bam_mod
        res = sharded.sharded_bam_walk(ctx, rank, world, rv, bam_mod._all_bytes(args.primary_bam), bam_mod._all_bytes(args.secondary_bam),
                                       mode=mode, score_src=_SCORE_SRC[tag_func], skip=skip, min_score=args.min_score, enabled_bins=enabled)
    else:
        # byte offset of each input's first record: the library's header pass on the raw bytes, on every rank
        with open(args.primary_sam.name, "rb") as fp, open(args.secondary_sam.name, "rb") as fs:
            hrc, _, rec_off, _, _ = _lib.process_headers(fp.fileno(), fs.fileno(), __version__)
        if hrc != _lib.XM_OK:
            raise IndexError('string index out of range')
        with open(args.primary_sam.name, "rb") as fp, open(args.secondary_sam.name, "rb") as fs:
            res = sharded.sharded_walk(ctx, rank, world, fp.fileno(), rec_off[0], fs.fileno(), rec_off[1], mode=mode, score_src=_SCORE_SRC[tag_func], skip=skip,
                                       min_score=args.min_score, enabled_bins=enabled)
    sharded.write_outputs(res, [f.fileno() if f else -1 for f in files], hdr_len)
    ctx.comm_barrier()
    if rank == 0:
        rv.cleanup()
        if res["status"] != _lib.XM_OK:
            ctx.close()
            raise RuntimeError("sharded walk failed ({0}): {1}".format(res["status"], res["message"]))

        class _R:
            counts = res["counts"]
        output_summary(_counter(_R, mode != _lib.MODE_SE))
    ctx.close()


if __name__ == '__main__':
    main()
