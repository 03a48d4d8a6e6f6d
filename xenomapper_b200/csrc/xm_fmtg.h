/*
 * xm_fmtg.h -- a 32-bit float as printf("%g") prints it, for the float aux values of BAM records.
 *
 * `samtools view` prints the BAM aux types f and B:f with "%g" (htslib sam.c: sam_format1), i.e. six significant digits,
 * fixed or exponent notation by the decimal exponent, trailing zeros dropped.  The rendering kernels do the same on the
 * device, digit for digit:
 *
 *   - the decimal exponent X from a table of powers of ten (a float is never within a double's rounding of a power of ten it
 *     does not equal);
 *   - the six digits D = round-half-even(v * 10^(5 - X)): one multiplication or one division by an EXACT power of ten in
 *     double precision when that power is exact (|5 - X| <= 22), two steps otherwise;
 *   - whenever the scaled value lies within 2^-20 of a half (where the double's own rounding could tip the decision) the
 *     comparison is redone exactly, with the float's integer mantissa, powers of five and shifts in a 256-bit integer.
 *
 * Plain host/device code; tests/emu runs it against the C library for every one of the 2^32 floats (tests/test_fmtg.py runs
 * a stride of them, scripts/check_fmtg_all.sh all).
 */
#pragma once
#include <stdint.h>
#include <string.h>

#include "xm_common.h"

namespace xm {

struct Big256 {
    uint32_t w[8];
};
XM_HD void big_set(Big256 &b, uint64_t v)
{
    for (int k = 0; k < 8; ++k) b.w[k] = 0;
    b.w[0] = (uint32_t)v; b.w[1] = (uint32_t)(v >> 32);
}
XM_HD void big_mul_small(Big256 &b, uint32_t m)
{
    uint64_t carry = 0;
    for (int k = 0; k < 8; ++k) { const uint64_t t = (uint64_t)b.w[k] * m + carry; b.w[k] = (uint32_t)t; carry = t >> 32; }
}
XM_HD void big_shl(Big256 &b, int n)
{
    const int ws = n >> 5, bs = n & 31;
    for (int k = 7; k >= 0; --k) {
        uint64_t v = 0;
        if (k - ws >= 0) v = (uint64_t)b.w[k - ws] << bs;
        if (bs && k - ws - 1 >= 0) v |= (uint64_t)b.w[k - ws - 1] >> (32 - bs);
        b.w[k] = (uint32_t)v;
    }
}
XM_HD int big_cmp(const Big256 &a, const Big256 &b)
{
    for (int k = 7; k >= 0; --k) if (a.w[k] != b.w[k]) return a.w[k] < b.w[k] ? -1 : 1;
    return 0;
}

XM_HD double fmtg_pow10(int k)          /* 10^k for 0 <= k <= 22: exact in a double */
{
    const double t[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    return t[k];
}
XM_HD double fmtg_mul(double a, double b)
{
#if XM_DEVICE_PASS
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
XM_HD double fmtg_div(double a, double b)
{
#if XM_DEVICE_PASS
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

/* 10^x, -50 <= x <= 45, within a few units in the last place (exact for 0 <= x <= 22) */
XM_HD double fmtg_p10(int x)
{
    double r = 1.0;
    for (int q = x; q > 0; q -= 22) r = fmtg_mul(r, fmtg_pow10(q > 22 ? 22 : q));
    for (int q = -x; q > 0; q -= 22) r = fmtg_div(r, fmtg_pow10(q > 22 ? 22 : q));
    return r;
}

/* exact sign of  mant * 2^e2 * 10^j - (2k + 1) / 2 :  -1, 0, +1   (is the scaled value below, at or above k + 1/2 ?) */
XM_HD int fmtg_exact_vs_half(uint32_t mant, int e2, int j, uint32_t k)
{
    /* compare  mant * 2^(e2 + 1) * 10^j  with  2k + 1;  10^j = 5^j 2^j; negative powers go to the other side */
    Big256 L, R;
    big_set(L, mant);
    big_set(R, 2ull * k + 1ull);
    int sh = e2 + 1 + j;                      /* power of two on the left */
    if (j >= 0) for (int q = 0; q < j; ++q) big_mul_small(L, 5u);
    else for (int q = 0; q < -j; ++q) big_mul_small(R, 5u);
    if (sh >= 0) big_shl(L, sh); else big_shl(R, -sh);
    return big_cmp(L, R);
}

/* f as "%g" prints it; o must hold 16 bytes.  Returns the length. */
XM_HD int fmt_g(float f, char *o)
{
    uint32_t bits;
    memcpy(&bits, &f, 4);
    int n = 0;
    if (bits >> 31) o[n++] = '-';
    const uint32_t ab = bits & 0x7fffffffu;
    if (ab == 0) { o[n++] = '0'; return n; }
    if (ab >= 0x7f800000u) {
        const char *t = ab == 0x7f800000u ? "inf" : "nan";
        for (int k = 0; k < 3; ++k) o[n++] = t[k];
        return n;
    }
    /* mant * 2^e2, exactly */
    uint32_t mant = ab & 0x7fffffu;
    int e2 = (int)(ab >> 23);
    if (e2 == 0) e2 = -149; else { mant |= 0x800000u; e2 -= 150; }
    float af;
    memcpy(&af, &ab, 4);
    const double v = (double)af;
    /* decimal exponent: 10^X <= v < 10^(X + 1) */
    int X = (int)(((e2 + 23) * 1233) >> 12);               /* floor(log10(2) * floor(log2 v)), off by one at most */
    if (e2 + 23 < 0) X = -(int)(((-(e2 + 23)) * 1233 + 4095) >> 12);
    /* the table values above 10^22 and the reciprocals are within an ulp or two of the power; a float is far from every power
     * of ten it does not equal (>= 1e-8 relative), and the powers it can equal (10^0 .. 10^10) are exact here */
    while (v < fmtg_p10(X)) --X;
    while (v >= fmtg_p10(X + 1)) ++X;
    /* six digits */
    const int j = 5 - X;                                    /* scale by 10^j */
    double s = v;
    for (int q = j; q > 0; q -= 22) s = fmtg_mul(s, fmtg_pow10(q > 22 ? 22 : q));
    for (int q = -j; q > 0; q -= 22) s = fmtg_div(s, fmtg_pow10(q > 22 ? 22 : q));
    uint32_t k = (uint32_t)s;                               /* floor */
    const double frac = s - (double)k;
    bool up;
    if (frac > 0.5 + 9.5e-7 || frac < 0.5 - 9.5e-7) up = frac > 0.5;
    else {
        const int c = fmtg_exact_vs_half(mant, e2, j, k);
        up = c > 0 || (c == 0 && (k & 1u));                  /* ties to even */
    }
    uint32_t D = k + (up ? 1u : 0u);
    /* s may have been rounded across an integer by the double arithmetic only where frac is near 0 or 1: D is then still the
     * nearest six-digit number (the error is far below one half).  A carry out of the sixth digit moves the exponent */
    if (D >= 1000000u) { D = 100000u; ++X; }
    if (D < 100000u) { D = 100000u; }                        /* cannot happen for X found above; keeps the digit loop in range */
    char dig[6];
    for (int q = 5; q >= 0; --q) { dig[q] = (char)('0' + D % 10u); D /= 10u; }
    int nd = 6;
    while (nd > 1 && dig[nd - 1] == '0') --nd;             /* trailing zeros go */
    if (X < -4 || X >= 6) {
        o[n++] = dig[0];
        if (nd > 1) { o[n++] = '.'; for (int q = 1; q < nd; ++q) o[n++] = dig[q]; }
        o[n++] = 'e';
        int ax = X;
        if (ax < 0) { o[n++] = '-'; ax = -ax; } else o[n++] = '+';
        o[n++] = (char)('0' + ax / 10); o[n++] = (char)('0' + ax % 10);
    } else if (X >= 0) {
        for (int q = 0; q <= X; ++q) o[n++] = q < nd ? dig[q] : '0';
        if (nd > X + 1) { o[n++] = '.'; for (int q = X + 1; q < nd; ++q) o[n++] = dig[q]; }
    } else {
        o[n++] = '0'; o[n++] = '.';
        for (int q = 0; q < -X - 1; ++q) o[n++] = '0';
        for (int q = 0; q < nd; ++q) o[n++] = dig[q];
    }
    return n;
}

}  // namespace xm
