// numparse.cuh -- one text field of a Hi-C matrix file -> IEEE binary64, correctly rounded, usable on host and device.
//
// The reference reads its input with bigmemory::read.big.matrix(type = 'double', sep = '\t') (R/TADpole.R:17), i.e.
// every field goes through a C decimal -> double conversion.  The same value must come out here, so the conversion
// is exact (round-to-nearest-even of the decimal value), not "close":
//   * <= 15 significant digits and |10-exponent| <= 22: one IEEE multiply or divide of two exactly representable
//     numbers (Clinger's fast path) -- every integer count takes this branch;
//   * otherwise the Eisel-Lemire algorithm on a 19-digit mantissa and 128-bit truncated powers of five
//     (pow5_table.inc); when more than 19 digits were given, the truncated mantissa w and w + 1 must round to the
//     same double, else the field is handed to the host (status NP_HOST), which runs strtod on it.
// Fields: optional blanks, sign, digits[.digits][e[+-]digits] | NA | NaN | Inf[inity]; an empty field is NA.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NP_HD __host__ __device__ __forceinline__
#else
#define NP_HD static inline
#endif

enum { NP_OK = 0, NP_HOST = 1 };

NP_HD double np_bits(uint64_t b) {
    union { uint64_t u; double d; } x;
    x.u = b;
    return x.d;
}

NP_HD uint64_t np_mulhi(uint64_t a, uint64_t b, uint64_t *lo) {
#ifdef __CUDA_ARCH__
    *lo = a * b;
    return __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    *lo = (uint64_t)p;
    return (uint64_t)(p >> 64);
#endif
}

NP_HD int np_clz(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

// w * 10^q -> bits of the nearest double (w != 0).  Returns false when the value cannot be decided here.
NP_HD bool np_eisel_lemire(uint64_t w, int64_t q, const uint64_t *pow5, uint64_t *bits_out) {
    if (q < -342) { *bits_out = 0; return true; }
    if (q > 308) { *bits_out = 0x7ff0000000000000ULL; return true; }
    const int lz = np_clz(w);
    w <<= lz;
    const int idx = 2 * (int)(q + 342);
    uint64_t lo, hi = np_mulhi(w, pow5[idx], &lo);
    if ((hi & 0x1ff) == 0x1ff) {               // the 55 leading bits could still change: refine with the low word
        uint64_t lo2, hi2 = np_mulhi(w, pow5[idx + 1], &lo2);
        lo += hi2;
        if (hi2 > lo) hi++;
    }
    const int upper = (int)(hi >> 63);
    uint64_t m = hi >> (upper + 64 - 52 - 3);
    int p2 = (int)((((int64_t)(152170 + 65536) * q) >> 16) + 63) + upper - lz + 1023;
    if (p2 <= 0) {                               // subnormal
        if (-p2 + 1 >= 64) { *bits_out = 0; return true; }
        m >>= -p2 + 1;
        m += (m & 1);
        m >>= 1;
        p2 = (m < (1ULL << 52)) ? 0 : 1;
        *bits_out = (m & ~(1ULL << 52)) | ((uint64_t)p2 << 52);
        return true;
    }
    if (lo <= 1 && q >= -4 && q <= 23 && (m & 3) == 1) {      // exactly half-way: round to even
        if ((m << (upper + 64 - 52 - 3)) == hi) m &= ~1ULL;
    }
    m += (m & 1);
    m >>= 1;
    if (m >= (2ULL << 52)) { m = 1ULL << 52; p2++; }
    m &= ~(1ULL << 52);
    if (p2 >= 0x7ff) { *bits_out = 0x7ff0000000000000ULL; return true; }
    *bits_out = m | ((uint64_t)p2 << 52);
    return true;
}

NP_HD bool np_isdigit(unsigned c) { return c - '0' < 10u; }
NP_HD unsigned np_lower(unsigned c) { return (c - 'A' < 26u) ? c + 32 : c; }

// p[0..len): the field without its separator.  pow10[0..22]: exact powers of ten; pow5: the 128-bit table.
NP_HD int np_parse_field(const unsigned char *p, int len, const double *pow10, const uint64_t *pow5, double *out) {
    int i = 0;
    while (i < len && (p[i] == ' ')) i++;
    while (len > i && (p[len - 1] == ' ' || p[len - 1] == '\r')) len--;
    if (i == len) { *out = np_bits(0x7ff8000000000000ULL); return NP_OK; }          // empty field: NA
    bool neg = false;
    if (p[i] == '-' || p[i] == '+') { neg = p[i] == '-'; i++; }
    if (i == len) return NP_HOST;
    const uint64_t sign = neg ? 0x8000000000000000ULL : 0;
    if (!np_isdigit(p[i]) && p[i] != '.') {
        const int rem = len - i;
        const unsigned a = np_lower(p[i]), b = rem > 1 ? np_lower(p[i + 1]) : 0, c = rem > 2 ? np_lower(p[i + 2]) : 0;
        if (a == 'n' && b == 'a' && (rem == 2 || (rem == 3 && c == 'n'))) { *out = np_bits(0x7ff8000000000000ULL); return NP_OK; }
        if (a == 'i' && b == 'n' && c == 'f' && rem == 3) { *out = np_bits(sign | 0x7ff0000000000000ULL); return NP_OK; }
        return NP_HOST;                                                                  // let strtod decide (or reject)
    }
    uint64_t w = 0;
    int nd = 0;
    int64_t e10 = 0;
    bool any = false, trunc = false;
    for (; i < len && np_isdigit(p[i]); i++) {
        const unsigned d = p[i] - '0';
        any = true;
        if (w == 0 && d == 0) continue;
        if (nd < 19) { w = w * 10 + d; nd++; }
        else { trunc |= d != 0; e10++; }
    }
    if (i < len && p[i] == '.') {
        i++;
        for (; i < len && np_isdigit(p[i]); i++) {
            const unsigned d = p[i] - '0';
            any = true;
            if (w == 0 && d == 0) { e10--; continue; }
            if (nd < 19) { w = w * 10 + d; nd++; e10--; }
            else trunc |= d != 0;
        }
    }
    if (!any) return NP_HOST;
    if (i < len && (p[i] == 'e' || p[i] == 'E')) {
        i++;
        bool eneg = false;
        if (i < len && (p[i] == '-' || p[i] == '+')) { eneg = p[i] == '-'; i++; }
        if (i == len || !np_isdigit(p[i])) return NP_HOST;
        int64_t ex = 0;
        for (; i < len && np_isdigit(p[i]); i++) if (ex < 100000) ex = ex * 10 + (p[i] - '0');
        e10 += eneg ? -ex : ex;
    }
    if (i != len) return NP_HOST;
    if (w == 0) { *out = np_bits(sign); return NP_OK; }
    if (!trunc && w <= (1ULL << 53) && e10 >= -22 && e10 <= 22) {
        double v = (double)w;
#ifdef __CUDA_ARCH__
        v = e10 < 0 ? __ddiv_rn(v, pow10[-e10]) : __dmul_rn(v, pow10[e10]);
#else
        v = e10 < 0 ? v / pow10[-e10] : v * pow10[e10];
#endif
        *out = neg ? -v : v;
        return NP_OK;
    }
    uint64_t b0;
    if (!np_eisel_lemire(w, e10, pow5, &b0)) return NP_HOST;
    if (trunc) {
        uint64_t b1;
        if (!np_eisel_lemire(w + 1, e10, pow5, &b1) || b1 != b0) return NP_HOST;
    }
    *out = np_bits(sign | b0);
    return NP_OK;
}
