/*
 * xm_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the xenomapper
 * read-binning walk, used as the parity checker for the CUDA path.
 *
 * Nothing under xenomapper_b200/ may link, import or call this file.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg use it, and there only as the checker or the timed CPU arm.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (1) the reference's own known answers (xenomapper/tests/test_xenomapper.py
 *       :93, :125, :158 SHA-224s; :165-184 decision table; :191-197, :203-209
 *       tag tables; :215-227, :232 CIGAR table), and
 *   (2) tests/golden/ *.json vectors produced by running the unmodified
 *       reference (tests/golden/make_golden.py imports /root/reference).
 *
 * Each function cites the lines of /root/reference/xenomapper/xenomapper.py
 * ("xm.py") whose behaviour it restates.  The restatement works on raw bytes
 * of the record region (header lines already removed) instead of Python text
 * streams; where Python's text layer matters (universal newlines, UTF-8
 * decoding, str.split() whitespace, float()/int() grammar) the rule is spelled
 * out next to the code.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "xm_oracle.h"

/* ------------------------------------------------------------------------ */
/* growable byte buffer for the six outputs                                  */

typedef struct {
    char *p;
    size_t n, cap;
} obuf;

static int obuf_put(obuf *b, const char *s, size_t n)
{
    if (b->n + n > b->cap) {
        size_t nc = b->cap ? b->cap * 2 : 1 << 16;
        while (nc < b->n + n) nc *= 2;
        char *q = (char *)realloc(b->p, nc);
        if (!q) return -1;
        b->p = q;
        b->cap = nc;
    }
    memcpy(b->p + b->n, s, n);
    b->n += n;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* text layer                                                                */

/* Python str.split() separators (Py_UNICODE_ISSPACE): ASCII 0x09-0x0d,
 * 0x1c-0x20, and the code points below.  xm.py:103 splits each line with a
 * bare .split(), so all of these separate tokens. */
static int is_py_space(uint32_t c)
{
    if (c <= 0x20) return (c >= 0x09 && c <= 0x0d) || (c >= 0x1c);
    if (c < 0x85) return 0;
    return c == 0x85 || c == 0xa0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200a) ||
           c == 0x2028 || c == 0x2029 || c == 0x202f || c == 0x205f || c == 0x3000;
}

/* Strict UTF-8 decode of one scalar at s[i..n).  Files are opened 'rt'
 * (xm.py:601, 605) so malformed input raises UnicodeDecodeError in the
 * reference.  Returns byte length, 0 on malformed input. */
static int utf8_next(const unsigned char *s, size_t i, size_t n, uint32_t *cp)
{
    unsigned char c = s[i];
    if (c < 0x80) { *cp = c; return 1; }
    if (c < 0xc2) return 0;
    if (c < 0xe0) {
        if (i + 1 >= n || (s[i + 1] & 0xc0) != 0x80) return 0;
        *cp = ((uint32_t)(c & 0x1f) << 6) | (s[i + 1] & 0x3f);
        return 2;
    }
    if (c < 0xf0) {
        if (i + 2 >= n || (s[i + 1] & 0xc0) != 0x80 || (s[i + 2] & 0xc0) != 0x80) return 0;
        uint32_t v = ((uint32_t)(c & 0x0f) << 12) | ((uint32_t)(s[i + 1] & 0x3f) << 6) | (s[i + 2] & 0x3f);
        if (v < 0x800 || (v >= 0xd800 && v <= 0xdfff)) return 0;
        *cp = v;
        return 3;
    }
    if (c < 0xf5) {
        if (i + 3 >= n || (s[i + 1] & 0xc0) != 0x80 || (s[i + 2] & 0xc0) != 0x80 || (s[i + 3] & 0xc0) != 0x80) return 0;
        uint32_t v = ((uint32_t)(c & 0x07) << 18) | ((uint32_t)(s[i + 1] & 0x3f) << 12) |
                     ((uint32_t)(s[i + 2] & 0x3f) << 6) | (s[i + 3] & 0x3f);
        if (v < 0x10000 || v > 0x10ffff) return 0;
        *cp = v;
        return 4;
    }
    return 0;
}

typedef struct {
    const unsigned char *s;
    size_t len;   /* token bytes */
} tok;

typedef struct {
    tok *t;
    size_t n, cap;
} toklist;

static int toklist_push(toklist *L, const unsigned char *s, size_t len)
{
    if (L->n == L->cap) {
        size_t nc = L->cap ? L->cap * 2 : 32;
        tok *q = (tok *)realloc(L->t, nc * sizeof(tok));
        if (!q) return -1;
        L->t = q;
        L->cap = nc;
    }
    L->t[L->n].s = s;
    L->t[L->n].len = len;
    L->n++;
    return 0;
}

/* One input stream: the reference's sam.readline() under universal newlines
 * ('\n', '\r\n' and a lone '\r' all end a line) followed by
 * .strip('\n').split()  (xm.py:103-104, 116-117). */
typedef struct {
    const unsigned char *buf;
    size_t len, pos;
    toklist cur;      /* tokens of the current line; n==0 means EOF or blank */
    int bad_utf8;
} stream;

/* returns 0 ok, -1 oom.  Leaves cur.n == 0 at EOF or on a blank line. */
static int stream_readline(stream *st)
{
    st->cur.n = 0;
    if (st->pos >= st->len) return 0;                 /* readline() == '' */
    const unsigned char *b = st->buf;
    size_t i = st->pos, e = i;
    while (e < st->len && b[e] != '\n' && b[e] != '\r') e++;
    size_t next = e;
    if (e < st->len) next = (b[e] == '\r' && e + 1 < st->len && b[e + 1] == '\n') ? e + 2 : e + 1;
    st->pos = next;
    /* split on whitespace runs */
    size_t tstart = (size_t)-1;
    while (i < e) {
        uint32_t cp;
        int k = utf8_next(b, i, e, &cp);
        if (!k) { st->bad_utf8 = 1; st->cur.n = 0; return 0; }
        if (is_py_space(cp)) {
            if (tstart != (size_t)-1) {
                if (toklist_push(&st->cur, b + tstart, i - tstart)) return -1;
                tstart = (size_t)-1;
            }
        } else if (tstart == (size_t)-1) {
            tstart = i;
        }
        i += (size_t)k;
    }
    if (tstart != (size_t)-1 && toklist_push(&st->cur, b + tstart, e - tstart)) return -1;
    return 0;
}

static int tok_eq(const tok *a, const tok *b)
{
    return a->len == b->len && memcmp(a->s, b->s, a->len) == 0;
}

/* a deep copy of a token list's (pointer,len) pairs: the pointers index the
 * caller's input buffer, which outlives the walk. */
static int toklist_copy(toklist *dst, const toklist *src)
{
    dst->n = 0;
    for (size_t i = 0; i < src->n; i++)
        if (toklist_push(dst, src->t[i].s, src->t[i].len)) return -1;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* number grammar                                                            */

/* Python's underscore rule (PEP 515): every '_' sits between two digits. */
static int strip_underscores(const unsigned char *s, size_t n, char *out, size_t outcap, size_t *outn)
{
    size_t k = 0;
    for (size_t i = 0; i < n; i++) {
        unsigned char c = s[i];
        if (c >= 0x80) return -2;                      /* non-ASCII digits: not restated */
        if (c == '_') {
            if (i == 0 || i + 1 >= n) return -1;
            if (!(s[i - 1] >= '0' && s[i - 1] <= '9') || !(s[i + 1] >= '0' && s[i + 1] <= '9')) return -1;
            continue;
        }
        if (k + 1 >= outcap) return -3;
        out[k++] = (char)c;
    }
    out[k] = 0;
    *outn = k;
    return 0;
}

static int ci_eq(const char *s, const char *lit)
{
    for (; *lit; s++, lit++) {
        char c = *s;
        if (c >= 'A' && c <= 'Z') c = (char)(c + 32);
        if (c != *lit) return 0;
    }
    return *s == 0;
}

/* float(text) as used by xm.py:191.  0 ok; XMO_ERR_VALUE on ValueError;
 * XMO_ERR_UNSUPPORTED for inputs the oracle does not restate. */
static int py_float(const unsigned char *s, size_t n, double *out)
{
    char tmp[512];
    size_t m;
    int r = strip_underscores(s, n, tmp, sizeof tmp, &m);
    if (r == -1) return XMO_ERR_VALUE;
    if (r < 0) return XMO_ERR_UNSUPPORTED;
    const char *p = tmp;
    int neg = 0;
    if (*p == '+' || *p == '-') { neg = (*p == '-'); p++; }
    if (ci_eq(p, "inf") || ci_eq(p, "infinity")) { *out = neg ? -INFINITY : INFINITY; return 0; }
    if (ci_eq(p, "nan")) { *out = NAN; return 0; }
    const char *q = p;
    int nd = 0;
    while (*q >= '0' && *q <= '9') { q++; nd++; }
    if (*q == '.') { q++; while (*q >= '0' && *q <= '9') { q++; nd++; } }
    if (!nd) return XMO_ERR_VALUE;
    if (*q == 'e' || *q == 'E') {
        q++;
        if (*q == '+' || *q == '-') q++;
        if (!(*q >= '0' && *q <= '9')) return XMO_ERR_VALUE;
        while (*q >= '0' && *q <= '9') q++;
    }
    if (*q) return XMO_ERR_VALUE;
    *out = strtod(tmp, NULL);   /* glibc strtod and CPython's dtoa are both correctly rounded */
    return 0;
}

/* int(text) as used by xm.py:250.  Values beyond int64 are not restated. */
static int py_int(const unsigned char *s, size_t n, int64_t *out)
{
    char tmp[512];
    size_t m;
    int r = strip_underscores(s, n, tmp, sizeof tmp, &m);
    if (r == -1) return XMO_ERR_VALUE;
    if (r < 0) return XMO_ERR_UNSUPPORTED;
    const char *p = tmp;
    int neg = 0;
    if (*p == '+' || *p == '-') { neg = (*p == '-'); p++; }
    if (!(*p >= '0' && *p <= '9')) return XMO_ERR_VALUE;
    uint64_t v = 0;
    for (; *p >= '0' && *p <= '9'; p++) {
        if (v > (UINT64_C(0x7fffffffffffffff) - (uint64_t)(*p - '0')) / 10) return XMO_ERR_UNSUPPORTED;
        v = v * 10 + (uint64_t)(*p - '0');
    }
    if (*p) return XMO_ERR_VALUE;
    *out = neg ? -(int64_t)v : (int64_t)v;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* tag extraction                                                            */

static int tok_contains2(const tok *t, char a, char b)
{
    for (size_t i = 0; i + 1 < t->len; i++)
        if (t->s[i] == (unsigned char)a && t->s[i + 1] == (unsigned char)b) return 1;
    return 0;
}

/* text after the last ':' of a token (xm.py:191 `split(':')[-1]`) */
static void after_last_colon(const tok *t, const unsigned char **s, size_t *n)
{
    size_t k = t->len;
    while (k > 0 && t->s[k - 1] != ':') k--;
    *s = t->s + k;
    *n = t->len - k;
}

/* get_tag, xm.py:176-191: substring match of the two-letter tag against every
 * token with index >= 11; none -> -inf; more than one -> ValueError; else
 * float() of the text after the last ':'. */
static int get_tag(const toklist *L, char a, char b, double *out)
{
    const tok *hit = NULL;
    int nhit = 0;
    for (size_t i = 11; i < L->n; i++)
        if (tok_contains2(&L->t[i], a, b)) { if (!nhit) hit = &L->t[i]; nhit++; }
    if (!nhit) { *out = -INFINITY; return 0; }
    if (nhit > 1) return XMO_ERR_VALUE;
    const unsigned char *s;
    size_t n;
    after_last_colon(hit, &s, &n);
    return py_float(s, n, out);
}

/* get_cigarbased_AS_tag with tag='AS', xm.py:247-256: first token >= 11 that
 * contains "NM" (no duplicate check), int() of its value, then every maximal
 * digit run in token 5 that is immediately followed by one of MIDNSHPX=. */
static int cigar_score(const toklist *L, double *out)
{
    const tok *nm = NULL;
    for (size_t i = 11; i < L->n && !nm; i++)
        if (tok_contains2(&L->t[i], 'N', 'M')) nm = &L->t[i];
    if (!nm) { *out = -INFINITY; return 0; }
    const unsigned char *s;
    size_t n;
    after_last_colon(nm, &s, &n);
    int64_t mm;
    int r = py_int(s, n, &mm);
    if (r) return r;
    const tok *cg = &L->t[5];
    int64_t n_id = 0, sum_id = 0, sum_s = 0;
    size_t i = 0;
    while (i < cg->len) {
        if (cg->s[i] < '0' || cg->s[i] > '9') { i++; continue; }
        uint64_t v = 0;
        int big = 0;
        while (i < cg->len && cg->s[i] >= '0' && cg->s[i] <= '9') {
            if (v > (UINT64_C(1) << 56)) big = 1;
            v = v * 10 + (uint64_t)(cg->s[i] - '0');
            i++;
        }
        if (i >= cg->len) break;
        unsigned char op = cg->s[i];
        if (op == 'I' || op == 'D') { if (big) return XMO_ERR_UNSUPPORTED; n_id++; sum_id += (int64_t)v; }
        else if (op == 'S') { if (big) return XMO_ERR_UNSUPPORTED; sum_s += (int64_t)v; }
    }
    if (mm > (INT64_C(1) << 56) || mm < -(INT64_C(1) << 56) || sum_id > (INT64_C(1) << 56) || sum_s > (INT64_C(1) << 56))
        return XMO_ERR_UNSUPPORTED;
    int64_t sc = -6 * mm - 5 * n_id - 3 * sum_id - 2 * sum_s;
    if (sc > (INT64_C(1) << 53) || sc < -(INT64_C(1) << 53)) return XMO_ERR_UNSUPPORTED;
    *out = (double)sc;
    return 0;
}

/* the tag_func seam (xm.py:299, 684-689): AS and XS of one record */
static int scores(const toklist *L, int score_src, double *as, double *xs)
{
    int r;
    if (score_src == XMO_SCORE_CIGAR_NM) {
        if ((r = cigar_score(L, as))) return r;         /* xm.py:247-256 */
        return get_tag(L, 'X', 'S', xs);                /* xm.py:245-246 */
    }
    if ((r = get_tag(L, 'A', 'S', as))) return r;
    if (score_src == XMO_SCORE_AS_ZS) return get_tag(L, 'Z', 'S', xs);   /* xm.py:204-206 */
    return get_tag(L, 'X', 'S', xs);
}

/* get_mapping_state, xm.py:275-289.  `not XS` is XS == 0.  Returns the state
 * index or -1 for the RuntimeError branch (NaN only). */
int xmo_mapping_state(double AS1, double XS1, double AS2, double XS2, double min_score)
{
    if (AS1 <= min_score && AS2 <= min_score) return XMO_UA;
    if (AS1 > min_score && (AS2 <= min_score || AS1 > AS2))
        return (XS1 == 0.0 || AS1 > XS1) ? XMO_PS : XMO_PM;
    if (AS1 == AS2) return XMO_UR;
    if (AS2 > min_score && (AS1 <= min_score || AS2 > AS1))
        return (XS2 == 0.0 || AS2 > XS2) ? XMO_SS : XMO_SM;
    return -1;
}

/* ------------------------------------------------------------------------ */
/* the walks                                                                 */

typedef struct {
    obuf out[6];
    unsigned enabled;
} sinks;

/* print('\t'.join(line), file=bin)  (xm.py:334 and siblings) */
static int emit(sinks *S, int bin, const toklist *L)
{
    if (!(S->enabled & (1u << bin))) return 0;
    for (size_t i = 0; i < L->n; i++) {
        if (i && obuf_put(&S->out[bin], "\t", 1)) return -1;
        if (obuf_put(&S->out[bin], (const char *)L->t[i].s, L->t[i].len)) return -1;
    }
    return obuf_put(&S->out[bin], "\n", 1);
}

/* liberal priority chain, xm.py:423-448: PS > SS > PM > SM > UR > UA */
static int liberal_bin(int f, int r)
{
    static const int order[6] = {XMO_PS, XMO_SS, XMO_PM, XMO_SM, XMO_UR, XMO_UA};
    for (int k = 0; k < 6; k++)
        if (f == order[k] || r == order[k]) return order[k];
    return -1;
}

/* conservative chain, xm.py:521-550 */
static int conservative_bin(int f, int r)
{
    if (f == XMO_UA || r == XMO_UA) return XMO_UA;
    int fp = (f == XMO_PS || f == XMO_PM), fs = (f == XMO_SS || f == XMO_SM);
    int rp = (r == XMO_PS || r == XMO_PM), rs = (r == XMO_SS || r == XMO_SM);
    if (f == XMO_UR || r == XMO_UR || (fp && rs) || (fs && rp)) return XMO_UR;
    if (f == XMO_PS || r == XMO_PS) return XMO_PS;
    if (f == XMO_SS || r == XMO_SS) return XMO_SS;
    if (f == XMO_PM || r == XMO_PM) return XMO_PM;
    return XMO_SM;
}

int xmo_pair_bin(int fwd, int rev, int conservative)
{
    return conservative ? conservative_bin(fwd, rev) : liberal_bin(fwd, rev);
}

static int fail(xmo_result *R, int code, uint64_t rec, const char *msg)
{
    R->err = code;
    R->err_record = rec;
    snprintf(R->errmsg, sizeof R->errmsg, "%s", msg);
    return code;
}

int xmo_classify(const void *prim, size_t prim_len, const void *sec, size_t sec_len,
                 const xmo_opts *o, xmo_result *R)
{
    memset(R, 0, sizeof *R);
    stream s1 = {(const unsigned char *)prim, prim_len, 0, {0, 0, 0}, 0};
    stream s2 = {(const unsigned char *)sec, sec_len, 0, {0, 0, 0}, 0};
    toklist prev1 = {0, 0, 0}, prev2 = {0, 0, 0}, name1 = {0, 0, 0}, name2 = {0, 0, 0};
    sinks S;
    memset(&S, 0, sizeof S);
    S.enabled = o->enabled_bins;
    int rc = 0;
    uint64_t idx = 0;        /* index in the yielded sequence */
    int have_prev = 0;

    if (stream_readline(&s1) || stream_readline(&s2)) { rc = fail(R, XMO_ERR_NOMEM, 0, "out of memory"); goto done; }

    /* getReadPairs, xm.py:105: stop at EOF or a blank line in either stream */
    while (s1.cur.n && s2.cur.n) {
        /* xm.py:106 (and again at :322/:399/:499) */
        if (!tok_eq(&s1.cur.t[0], &s2.cur.t[0])) { rc = fail(R, XMO_ERR_ASSERT, idx, "QNAME mismatch"); goto done; }

        if (o->mode == XMO_MODE_SE) {
            /* main_single_end, xm.py:321-351 */
            double a1, x1, a2, x2;
            int r;
            if ((r = scores(&s1.cur, o->score_src, &a1, &x1)) || (r = scores(&s2.cur, o->score_src, &a2, &x2))) {
                rc = fail(R, r, idx, "tag value"); goto done;
            }
            int st = xmo_mapping_state(a1, x1, a2, x2, o->min_score);
            if (st < 0) { rc = fail(R, XMO_ERR_RUNTIME, idx, "processing logic"); goto done; }
            R->counts[st]++;
            int e = 0;
            switch (st) {
            case XMO_PS: case XMO_PM: case XMO_UA: e = emit(&S, st, &s1.cur); break;
            case XMO_SS: case XMO_SM: e = emit(&S, st, &s2.cur); break;
            default: e = emit(&S, XMO_UR, &s1.cur) || emit(&S, XMO_UR, &s2.cur); break;
            }
            if (e) { rc = fail(R, XMO_ERR_NOMEM, idx, "out of memory"); goto done; }
        } else {
            /* main_paired_end / conservative_main_paired_end, xm.py:398-452, 498-554.
             * A unit fires when the previous yielded primary QNAME equals this one. */
            if (have_prev && tok_eq(&prev1.t[0], &s1.cur.t[0])) {
                double pa1, px1, pa2, px2, a1, x1, a2, x2;
                int r;
                if ((r = scores(&prev1, o->score_src, &pa1, &px1)) || (r = scores(&prev2, o->score_src, &pa2, &px2)) ||
                    (r = scores(&s1.cur, o->score_src, &a1, &x1)) || (r = scores(&s2.cur, o->score_src, &a2, &x2))) {
                    rc = fail(R, r, idx, "tag value"); goto done;
                }
                int f = xmo_mapping_state(pa1, px1, pa2, px2, o->min_score);
                if (f < 0) { rc = fail(R, XMO_ERR_RUNTIME, idx, "processing logic"); goto done; }
                int v = xmo_mapping_state(a1, x1, a2, x2, o->min_score);
                if (v < 0) { rc = fail(R, XMO_ERR_RUNTIME, idx, "processing logic"); goto done; }
                R->counts[f * 6 + v]++;
                int bin = xmo_pair_bin(f, v, o->mode == XMO_MODE_PE_CONSERVATIVE);
                int e = 0;
                if (bin == XMO_PS || bin == XMO_PM || bin == XMO_UA)
                    e = emit(&S, bin, &prev1) || emit(&S, bin, &s1.cur);
                else if (bin == XMO_SS || bin == XMO_SM)
                    e = emit(&S, bin, &prev2) || emit(&S, bin, &s2.cur);
                else
                    e = emit(&S, bin, &prev1) || emit(&S, bin, &s1.cur) || emit(&S, bin, &prev2) || emit(&S, bin, &s2.cur);
                if (e) { rc = fail(R, XMO_ERR_NOMEM, idx, "out of memory"); goto done; }
            }
            if (toklist_copy(&prev1, &s1.cur) || toklist_copy(&prev2, &s2.cur)) { rc = fail(R, XMO_ERR_NOMEM, idx, "out of memory"); goto done; }
            have_prev = 1;
        }
        idx++;

        /* advance, xm.py:108-117 */
        if (o->skip_repeated) {
            if (toklist_copy(&name1, &s1.cur) || toklist_copy(&name2, &s2.cur)) { rc = fail(R, XMO_ERR_NOMEM, idx, "out of memory"); goto done; }
            while (s1.cur.n && s2.cur.n && tok_eq(&s1.cur.t[0], &name1.t[0]))
                if (stream_readline(&s1)) { rc = fail(R, XMO_ERR_NOMEM, idx, "out of memory"); goto done; }
            while (s1.cur.n && s2.cur.n && tok_eq(&s2.cur.t[0], &name2.t[0]))
                if (stream_readline(&s2)) { rc = fail(R, XMO_ERR_NOMEM, idx, "out of memory"); goto done; }
        } else {
            if (stream_readline(&s1) || stream_readline(&s2)) { rc = fail(R, XMO_ERR_NOMEM, idx, "out of memory"); goto done; }
        }
    }
    if (s1.bad_utf8 || s2.bad_utf8) rc = fail(R, XMO_ERR_UNICODE, idx, "invalid UTF-8");

done:
    R->n_yielded = idx;
    for (int b = 0; b < 6; b++) { R->data[b] = S.out[b].p; R->len[b] = S.out[b].n; }
    free(s1.cur.t); free(s2.cur.t); free(prev1.t); free(prev2.t); free(name1.t); free(name2.t);
    return rc;
}

void xmo_free(xmo_result *R)
{
    for (int b = 0; b < 6; b++) { free(R->data[b]); R->data[b] = NULL; R->len[b] = 0; }
}

/* convenience for tests: scores of a single line given as raw bytes */
int xmo_line_scores(const void *line, size_t len, int score_src, double *as, double *xs)
{
    stream st = {(const unsigned char *)line, len, 0, {0, 0, 0}, 0};
    if (stream_readline(&st)) return XMO_ERR_NOMEM;
    int r = st.bad_utf8 ? XMO_ERR_UNICODE : scores(&st.cur, score_src, as, xs);
    free(st.cur.t);
    return r;
}
