"""SAM text -> BAM bytes, for tests only (SAM/BAM specification sections 4.1-4.2).

Builds the inputs of the BAM tests from the synthetic SAM pairs: rendering the result on the GPU must give the
SAM text back byte for byte.  Integer tags take the smallest type that holds the value, as samtools does.
"""
import struct
import zlib

_CIG = {c: i for i, c in enumerate("MIDNSHP=X")}
_NT = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def _aux(tok):
    tag, typ, val = tok.split(":", 2)
    t = tag.encode()
    if typ == "i":
        v = int(val)
        for code, lo, hi, fmt in (("C", 0, 255, "<B"), ("c", -128, 127, "<b"), ("S", 0, 65535, "<H"), ("s", -32768, 32767, "<h"),
                                  ("I", 0, 2 ** 32 - 1, "<I"), ("i", -2 ** 31, 2 ** 31 - 1, "<i")):
            if lo <= v <= hi:
                return t + code.encode() + struct.pack(fmt, v)
        raise ValueError(val)
    if typ == "A":
        return t + b"A" + val.encode()
    if typ in "ZH":
        return t + typ.encode() + val.encode() + b"\0"
    if typ == "B":
        sub, *vals = val.split(",")
        fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I"}[sub]
        return t + b"B" + sub.encode() + struct.pack("<I", len(vals)) + b"".join(struct.pack("<" + fmt, int(v)) for v in vals)
    raise ValueError(typ)


def _reg2bin(beg, end):
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def record(line, ref_ids):
    f = line.split("\t")
    qname, flag, rname, pos, mapq, cigar, rnext, pnext, tlen, seq, qual = f[:11]
    name = (qname if qname != "*" else "").encode() + b"\0"
    ops = []
    if cigar != "*":
        n = ""
        for ch in cigar:
            if ch.isdigit():
                n += ch
            else:
                ops.append((int(n) << 4) | _CIG[ch])
                n = ""
    l_seq = 0 if seq == "*" else len(seq)
    packed = bytearray((l_seq + 1) // 2)
    for k in range(l_seq):
        packed[k >> 1] |= _NT[seq[k]] << (0 if k & 1 else 4)
    q = bytes([0xff] * l_seq) if qual == "*" else bytes(ord(c) - 33 for c in qual)
    ref = ref_ids.get(rname, -1)
    nref = ref if rnext == "=" else ref_ids.get(rnext, -1)
    p0 = int(pos) - 1
    reflen = sum(v >> 4 for v in ops if (v & 15) in (0, 2, 3, 7, 8)) or 1
    body = struct.pack("<iiBBHHHiiii", ref, p0, len(name), int(mapq), _reg2bin(max(p0, 0), max(p0, 0) + reflen), len(ops), int(flag),
                       l_seq, nref, int(pnext) - 1, int(tlen))
    body += name + b"".join(struct.pack("<I", v) for v in ops) + bytes(packed) + q + b"".join(_aux(t) for t in f[11:])
    return struct.pack("<i", len(body)) + body


def bgzf(data, block=0xff00, level=1):
    out = bytearray()
    for k in list(range(0, len(data), block)) + [len(data)]:
        chunk = data[k:k + block] if k < len(data) else b""
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        out += struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(comp) + 25)
        out += comp + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
    return bytes(out)


def sam_to_bam(header_text, record_text, block=0xff00, level=1):
    """header_text: '@' lines; record_text: SAM records (bytes or str).  Returns BAM file bytes."""
    if isinstance(header_text, bytes):
        header_text = header_text.decode()
    if isinstance(record_text, bytes):
        record_text = record_text.decode()
    refs = []
    for line in header_text.split("\n"):
        if line.startswith("@SQ"):
            d = dict(t.split(":", 1) for t in line.split("\t")[1:])
            refs.append((d["SN"], int(d["LN"])))
    ref_ids = {n: i for i, (n, _) in enumerate(refs)}
    text = header_text.encode()
    raw = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs)))
    for n, ln in refs:
        raw += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", ln)
    for line in record_text.split("\n"):
        if line:
            raw += record(line, ref_ids)
    return bgzf(bytes(raw), block, level)
