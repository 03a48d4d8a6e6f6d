/*
 * xm_synth.c -- deterministic synthetic SAM pair generator (host, plain C).
 *
 * Produces the primary-species and secondary-species SAM record regions the
 * bench and the parity tests run on: Bowtie2 --local-style single-end 150 bp
 * records (BASELINE.json configs[1]), interlaced 2x150 bp pairs (configs[2])
 * and HISAT-style records with ZS:i and XS:A (configs[3]).  Every record is a
 * pure function of (seed, style, record index), so any sub-range -- e.g. one
 * GPU's shard -- can be generated on its own and concatenates to the same
 * bytes.
 *
 * Distributions follow SURVEY.md section 8(d): mapped fraction 0.98 primary /
 * 0.16 secondary, XS on ~0.42 of mapped reads, XS==AS on ~0.14 of those, and
 * a latent category mix close to the published SRR879369 table
 * (PS 91.8 %, PM 6.4 %, UA 1.25 %, SS 0.47 %, SM 0.046 %, UR 0.028 %).
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define XS_STYLE_SE_BOWTIE2 0
#define XS_STYLE_PE_BOWTIE2 1
#define XS_STYLE_PE_HISAT   2

typedef struct { uint64_t s; } rng_t;

static inline uint64_t mix64(uint64_t z)
{
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static inline uint64_t rnd(rng_t *r) { r->s += 0x9e3779b97f4a7c15ULL; return mix64(r->s); }
static inline uint32_t rnd_below(rng_t *r, uint32_t n) { return (uint32_t)(((rnd(r) >> 32) * (uint64_t)n) >> 32); }
static inline double rnd_unit(rng_t *r) { return (double)(rnd(r) >> 11) * (1.0 / 9007199254740992.0); }

static inline char *put_u(char *p, uint64_t v)
{
    char t[24];
    int n = 0;
    do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = t[--n];
    return p;
}
static inline char *put_i(char *p, int64_t v)
{
    if (v < 0) { *p++ = '-'; return put_u(p, (uint64_t)(-v)); }
    return put_u(p, (uint64_t)v);
}
static inline char *put_s(char *p, const char *s) { while (*s) *p++ = *s++; return p; }

enum { C_PS, C_PM, C_UA, C_SS, C_SM, C_UR };

static int draw_category(rng_t *r)
{
    double u = rnd_unit(r);
    if (u < 0.91800) return C_PS;
    if (u < 0.98200) return C_PM;
    if (u < 0.99450) return C_UA;
    if (u < 0.99920) return C_SS;
    if (u < 0.99966) return C_SM;
    return C_UR;
}

typedef struct {
    int mapped;
    int as, xs;         /* xs < -1000000 means absent */
    int rev;            /* strand */
} aln_t;

#define XS_ABSENT (-2000000000)

/* scores of one read in both species, given its latent category */
static void draw_scores(rng_t *r, int cat, int hisat, aln_t *p, aln_t *s)
{
    /* a "good" local score skewed to the 2*len maximum; hisat: 0 is perfect */
    int hi;
    {
        double u = rnd_unit(r);
        int drop = (u < 0.55) ? 0 : (int)(240.0 * (u - 0.55) * (u - 0.55) / (0.45 * 0.45));
        hi = 300 - drop;
        if (hi < 62) hi = 62;
    }
    int lo = 60 + (int)rnd_below(r, (uint32_t)(hi - 60));   /* strictly below hi */
    p->rev = (int)(rnd(r) & 1); s->rev = (int)(rnd(r) & 1);
    p->xs = s->xs = XS_ABSENT;
    int other_mapped = rnd_unit(r) < 0.155;
    switch (cat) {
    case C_PS: case C_PM:
        p->mapped = 1; p->as = hi;
        s->mapped = other_mapped; s->as = lo;
        if (cat == C_PM) p->xs = hi;
        else if (rnd_unit(r) < 0.38) p->xs = 20 + (int)rnd_below(r, (uint32_t)(hi - 20));
        if (s->mapped && rnd_unit(r) < 0.42) s->xs = (rnd_unit(r) < 0.14) ? lo : 20 + (int)rnd_below(r, (uint32_t)(lo - 19));
        break;
    case C_SS: case C_SM:
        s->mapped = 1; s->as = hi;
        p->mapped = rnd_unit(r) < 0.6; p->as = lo;
        if (cat == C_SM) s->xs = hi;
        else if (rnd_unit(r) < 0.38) s->xs = 20 + (int)rnd_below(r, (uint32_t)(hi - 20));
        if (p->mapped && rnd_unit(r) < 0.42) p->xs = (rnd_unit(r) < 0.14) ? lo : 20 + (int)rnd_below(r, (uint32_t)(lo - 19));
        break;
    case C_UR:
        p->mapped = s->mapped = 1; p->as = s->as = hi;
        if (rnd_unit(r) < 0.42) p->xs = 20 + (int)rnd_below(r, (uint32_t)(hi - 19));
        if (rnd_unit(r) < 0.42) s->xs = 20 + (int)rnd_below(r, (uint32_t)(hi - 19));
        break;
    default:
        p->mapped = s->mapped = 0; p->as = s->as = 0;
        break;
    }
    if (hisat) {
        /* HISAT: AS <= 0, 0 is a perfect hit; ZS (next best) <= AS.  Map the
         * local score s to -(300-s)/2 so perfect reads carry AS:i:0 and the
         * `not XS` quirk (xm.py:278) is exercised by ZS:i:0. */
        if (p->mapped) { p->as = -(300 - p->as) / 2; if (p->xs != XS_ABSENT) p->xs = -(300 - p->xs) / 2; }
        if (s->mapped) { s->as = -(300 - s->as) / 2; if (s->xs != XS_ABSENT) s->xs = -(300 - s->xs) / 2; }
    }
}

static const char BASES[4] = {'A', 'C', 'G', 'T'};

#define READ_LEN 150

typedef struct {
    char seq[READ_LEN], qual[READ_LEN];
} read_t;

static void draw_read(rng_t *r, read_t *rd)
{
    for (int i = 0; i < READ_LEN; i += 32) {
        uint64_t w = rnd(r);
        for (int j = 0; j < 32 && i + j < READ_LEN; j++, w >>= 2) rd->seq[i + j] = BASES[w & 3];
    }
    if ((rnd(r) & 63) == 0) rd->seq[rnd_below(r, READ_LEN)] = 'N';
    /* Illumina 1.8+ style: plateau of J/F/A with a decaying tail and '#' run */
    static const char QS[8] = {'J', 'J', 'J', 'F', 'F', 'A', '<', '7'};
    int tail = READ_LEN - (int)rnd_below(r, 40);
    for (int i = 0; i < READ_LEN; i += 21) {
        uint64_t w = rnd(r);
        for (int j = 0; j < 21 && i + j < READ_LEN; j++, w >>= 3) {
            int k = (int)(w & 7);
            rd->qual[i + j] = (i + j < tail) ? QS[k] : (k < 3 ? '#' : QS[4 + (k & 3)]);
        }
    }
    rd->qual[0] = 'A'; rd->qual[1] = 'A';
}

static char *put_seq(char *p, const read_t *rd, int rev)
{
    if (!rev) { memcpy(p, rd->seq, READ_LEN); p += READ_LEN; *p++ = '\t'; memcpy(p, rd->qual, READ_LEN); return p + READ_LEN; }
    for (int i = READ_LEN - 1; i >= 0; i--) {
        char c = rd->seq[i];
        *p++ = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    }
    *p++ = '\t';
    for (int i = READ_LEN - 1; i >= 0; i--) *p++ = rd->qual[i];
    return p;
}

static const char *CHR_P[8] = {"chr1", "chr2", "chr7", "chr11", "chr17", "chrX", "chrM", "chr19"};
static const char *CHR_S[8] = {"1", "2", "7", "11", "17", "X", "MT", "19"};

/* one aligned-or-not record; returns new write pointer */
static char *put_record(char *p, rng_t *r, const char *qname, int qlen, const read_t *rd, const aln_t *a,
                        int secondary, int style, int mate, const aln_t *mate_aln)
{
    int paired = style != XS_STYLE_SE_BOWTIE2;
    int hisat = style == XS_STYLE_PE_HISAT;
    memcpy(p, qname, (size_t)qlen); p += qlen; *p++ = '\t';
    int flag = 0;
    if (paired) {
        flag |= 1 | (mate ? 128 : 64);
        if (!a->mapped) flag |= 4;
        if (!mate_aln->mapped) flag |= 8;
        if (a->mapped && a->rev) flag |= 16;
        if (mate_aln->mapped && mate_aln->rev) flag |= 32;
        if (a->mapped && mate_aln->mapped && a->rev != mate_aln->rev) flag |= 2;
    } else {
        flag = a->mapped ? (a->rev ? 16 : 0) : 4;
    }
    p = put_u(p, (uint64_t)flag); *p++ = '\t';
    if (!a->mapped) {
        p = put_s(p, "*\t0\t0\t*\t*\t0\t0\t");
        p = put_seq(p, rd, 0);
        if (paired) { p = put_s(p, mate_aln->mapped ? "\tYS:i:" : ""); if (mate_aln->mapped) p = put_i(p, mate_aln->as); }
        p = put_s(p, paired ? "\tYT:Z:UP" : "\tYT:Z:UU");
        if (hisat) p = put_s(p, "\tYF:Z:NS");
        *p++ = '\n';
        return p;
    }
    const char *chr = (secondary ? CHR_S : CHR_P)[rnd_below(r, 8)];
    uint32_t pos = 1 + rnd_below(r, 150000000u);
    p = put_s(p, chr); *p++ = '\t';
    p = put_u(p, pos); *p++ = '\t';
    int mapq = (a->xs == a->as) ? (int)rnd_below(r, 2) : (a->xs == XS_ABSENT ? 42 + (int)rnd_below(r, 3) : (int)rnd_below(r, 42));
    p = put_u(p, (uint64_t)mapq); *p++ = '\t';
    /* CIGAR: 70 % full match, 25 % soft clipped, 5 % with one insertion or deletion */
    double u = rnd_unit(r);
    int nm = 0, xo = 0, xg = 0, md_del = 0;
    if (u < 0.70) { p = put_s(p, "150M"); }
    else if (u < 0.95) {
        int l = (int)rnd_below(r, 30), t = (int)rnd_below(r, 30);
        if (!l && !t) l = 1 + (int)rnd_below(r, 20);
        if (l) { p = put_u(p, (uint64_t)l); *p++ = 'S'; }
        p = put_u(p, (uint64_t)(READ_LEN - l - t)); *p++ = 'M';
        if (t) { p = put_u(p, (uint64_t)t); *p++ = 'S'; }
    } else {
        int a1 = 10 + (int)rnd_below(r, 100), g = 1 + (int)rnd_below(r, 3), ins = (int)(rnd(r) & 1);
        p = put_u(p, (uint64_t)a1); *p++ = 'M';
        p = put_u(p, (uint64_t)g); *p++ = ins ? 'I' : 'D';
        p = put_u(p, (uint64_t)(READ_LEN - a1 - (ins ? g : 0))); *p++ = 'M';
        xo = 1; xg = g; nm += g; md_del = ins ? 0 : g;
    }
    *p++ = '\t';
    if (paired && mate_aln->mapped) {
        int64_t d = (int64_t)rnd_below(r, 400) + READ_LEN;
        p = put_s(p, "=\t"); p = put_u(p, pos + (uint64_t)(a->rev ? 0 : d - READ_LEN)); *p++ = '\t';
        p = put_i(p, a->rev ? -d : d); *p++ = '\t';
    } else {
        p = put_s(p, "*\t0\t0\t");
    }
    p = put_seq(p, rd, a->rev);
    int xm = (int)rnd_below(r, 4) * (u < 0.5 ? 0 : 1);
    nm += xm;
    p = put_s(p, "\tAS:i:"); p = put_i(p, a->as);
    if (a->xs != XS_ABSENT) { p = put_s(p, hisat ? "\tZS:i:" : "\tXS:i:"); p = put_i(p, a->xs); }
    p = put_s(p, "\tXN:i:0\tXM:i:"); p = put_u(p, (uint64_t)xm);
    p = put_s(p, "\tXO:i:"); p = put_u(p, (uint64_t)xo);
    p = put_s(p, "\tXG:i:"); p = put_u(p, (uint64_t)xg);
    p = put_s(p, "\tNM:i:"); p = put_u(p, (uint64_t)nm);
    p = put_s(p, "\tMD:Z:");
    if (xm) { int a1 = 1 + (int)rnd_below(r, 140); p = put_u(p, (uint64_t)a1); *p++ = BASES[rnd(r) & 3]; p = put_u(p, (uint64_t)(148 - a1)); }
    else if (md_del) { p = put_s(p, "70^"); for (int i = 0; i < md_del; i++) *p++ = BASES[rnd(r) & 3]; p = put_s(p, "80"); }
    else p = put_s(p, "150");
    if (paired) {
        if (mate_aln->mapped) { p = put_s(p, "\tYS:i:"); p = put_i(p, mate_aln->as); }
        p = put_s(p, mate_aln->mapped ? ((flag & 2) ? "\tYT:Z:CP" : "\tYT:Z:DP") : "\tYT:Z:UP");
    } else {
        p = put_s(p, "\tYT:Z:UU");
    }
    if (hisat) {
        p = put_s(p, (rnd(r) & 1) ? "\tXS:A:+" : "\tXS:A:-");
        p = put_s(p, "\tNH:i:"); p = put_u(p, (uint64_t)(a->xs == a->as ? 2 + rnd_below(r, 4) : 1));
    }
    *p++ = '\n';
    return p;
}

#define XS_MAX_RECORD 640

/* Generate records [first, first+count) of both streams.  Buffers must hold
 * count * XS_MAX_RECORD bytes.  Returns 0, lengths in *len_p / *len_s. */
int xm_synth_generate(uint64_t seed, int style, uint64_t first, uint64_t count,
                      char *out_p, uint64_t cap_p, uint64_t *len_p,
                      char *out_s, uint64_t cap_s, uint64_t *len_s)
{
    if (cap_p < count * XS_MAX_RECORD || cap_s < count * XS_MAX_RECORD) return -1;
    char *pp = out_p, *ps = out_s;
    int paired = style != XS_STYLE_SE_BOWTIE2;
    int hisat = style == XS_STYLE_PE_HISAT;
    for (uint64_t rec = first; rec < first + count; rec++) {
        uint64_t unit = paired ? rec >> 1 : rec;      /* mates share the unit stream for QNAME + pair category */
        int mate = paired ? (int)(rec & 1) : 0;
        rng_t ru = {mix64(seed ^ (unit * 0xd1342543de82ef95ULL))};
        char qname[64];
        char *q = qname;
        q = put_s(q, "SIM-B200:"); q = put_u(q, 60 + (unit >> 22) % 40);
        q = put_s(q, ":H7T2NDSXX:"); q = put_u(q, 1 + (unit >> 20) % 4); *q++ = ':';
        q = put_u(q, 1101 + (unit >> 12) % 256 * 7 % 1600); *q++ = ':';
        q = put_u(q, 1000 + rnd_below(&ru, 31000)); *q++ = ':';
        q = put_u(q, 1000 + (unit & 0xfff) * 8 + rnd_below(&ru, 8));
        int qlen = (int)(q - qname);
        int pair_cat = draw_category(&ru);
        aln_t ap[2], as_[2];
        read_t rd;
        for (int m = 0; m <= (paired ? 1 : 0); m++) {
            /* a mate usually follows the pair's category; 3 % go their own way */
            int cat = pair_cat;
            if (paired && rnd_unit(&ru) < 0.03) cat = draw_category(&ru);
            draw_scores(&ru, cat, hisat, &ap[m], &as_[m]);
        }
        if (!paired) { ap[1] = ap[0]; as_[1] = as_[0]; }
        rng_t rr = {mix64(seed ^ (rec * 0x9e3779b97f4a7c15ULL) ^ 0x5851f42d4c957f2dULL)};
        draw_read(&rr, &rd);
        rng_t r1 = {rnd(&rr)}, r2 = {rnd(&rr)};
        pp = put_record(pp, &r1, qname, qlen, &rd, &ap[mate], 0, style, mate, &ap[mate ^ 1]);
        ps = put_record(ps, &r2, qname, qlen, &rd, &as_[mate], 1, style, mate, &as_[mate ^ 1]);
    }
    *len_p = (uint64_t)(pp - out_p);
    *len_s = (uint64_t)(ps - out_s);
    return 0;
}

uint64_t xm_synth_max_record(void) { return XS_MAX_RECORD; }
