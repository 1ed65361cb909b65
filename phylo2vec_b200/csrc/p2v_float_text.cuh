// float64 <-> decimal text, exact, usable on the device and on the host.
//
//   f64_to_shortest(x, buf): the shortest decimal string that parses back to x, laid out like the Rust `ryu`
//       crate's Buffer::format (what phylo2vec/src/matrix/convert.rs:37-57 writes into Newick strings):
//       "0.1", "3.0", "1e-7", "1.5e300", "0.00001", "NaN", "inf", "-inf".
//       Algorithm: Ryu (Adams, "Ryu: fast float-to-string conversion", PLDI 2018), full 128-bit tables.
//   parse_f64(s, len, &x): correctly rounded value of a decimal literal with Rust's `str::parse::<f64>` grammar
//       (newick/mod.rs:189-251 parses branch lengths with it): [+-] digits [. digits] [e [+-] digits] with at
//       least one mantissa digit, or inf / infinity / nan in any case.
//       Algorithm: Eisel-Lemire (Lemire, "Number parsing at a gigabyte per second", SP&E 2021; with the
//       128-bit product it needs no fallback for up to 19 significant digits, Mushtak & Lemire 2023).
//       Longer mantissas are cut to 19 digits and decided when the cut and the cut + 1 round to the same
//       double (virtually always); otherwise PARSE_F64_UNDECIDED is returned and the caller takes the literal to
//       parse_f64_exact (big-number comparison with the midpoint of the two candidates).
//
// Compiled as plain C++ too (tests/float_text_host.cpp checks both routines against Python's repr / float on
// millions of values without a GPU).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define P2V_HD __host__ __device__ __forceinline__
#else
#define P2V_HD static inline
#endif

namespace p2v_ft {

#if defined(__CUDACC__)
// one copy of the tables in device memory and one on the host
namespace dev {
#define P2V_TABLE_QUAL static __device__
#include "p2v_float_tables.h"
}  // namespace dev
namespace host {
#undef P2V_TABLE_QUAL
#define P2V_TABLE_QUAL static
#include "p2v_float_tables.h"
}  // namespace host
#if defined(__CUDA_ARCH__)
#define P2V_FT_TAB(name) dev::name
#else
#define P2V_FT_TAB(name) host::name
#endif
#else
namespace host {
#define P2V_TABLE_QUAL static
#include "p2v_float_tables.h"
}
#define P2V_FT_TAB(name) host::name
#endif

P2V_HD uint64_t umulh(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
P2V_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}
P2V_HD uint64_t f64_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
P2V_HD double bits_f64(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

// ------------------------------------------------------------------------------------------ Ryu
P2V_HD uint32_t ryu_pow5bits(int32_t e) { return (uint32_t)(((uint32_t)e * 1217359u) >> 19) + 1u; }
P2V_HD uint32_t ryu_log10pow2(int32_t e) { return ((uint32_t)e * 78913u) >> 18; }
P2V_HD uint32_t ryu_log10pow5(int32_t e) { return ((uint32_t)e * 732923u) >> 20; }
P2V_HD uint32_t ryu_pow5factor(uint64_t v) {
    uint32_t c = 0;
    for (;;) {
        const uint64_t q = v / 5;
        if (v - 5 * q != 0) break;
        v = q; ++c;
    }
    return c;
}
// (m * mul) >> j for a 128-bit mul = {lo, hi}, 64 <= j < 128 + 64
P2V_HD uint64_t ryu_mulshift(uint64_t m, const uint64_t* mul, int32_t j) {
    const uint64_t b0_hi = umulh(m, mul[0]);
    const uint64_t b2_lo = m * mul[1], b2_hi = umulh(m, mul[1]);
    const uint64_t sum_lo = b0_hi + b2_lo;
    const uint64_t sum_hi = b2_hi + (sum_lo < b0_hi ? 1u : 0u);
    const int32_t s = j - 64;                         // 0 < s < 64 for every double
    return (sum_lo >> s) | (sum_hi << (64 - s));
}

struct Decimal64 { uint64_t mantissa; int32_t exponent; };

// shortest (mantissa, exponent) with mantissa * 10^exponent == x for a finite, non-zero double given by its fields
P2V_HD Decimal64 ryu_d2d(uint64_t ieee_mantissa, uint32_t ieee_exponent) {
    int32_t e2;
    uint64_t m2;
    if (ieee_exponent == 0) { e2 = 1 - 1023 - 52 - 2; m2 = ieee_mantissa; }
    else { e2 = (int32_t)ieee_exponent - 1023 - 52 - 2; m2 = (1ull << 52) | ieee_mantissa; }
    const bool accept = (m2 & 1) == 0;
    const uint64_t mv = 4 * m2;
    const uint32_t mm_shift = (ieee_mantissa != 0 || ieee_exponent <= 1) ? 1u : 0u;
    uint64_t vr, vp, vm;
    int32_t e10;
    bool vm_tz = false, vr_tz = false;
    if (e2 >= 0) {
        const uint32_t q = ryu_log10pow2(e2) - (e2 > 3 ? 1u : 0u);
        e10 = (int32_t)q;
        const int32_t k = 125 + (int32_t)ryu_pow5bits((int32_t)q) - 1;
        const int32_t i = -e2 + (int32_t)q + k;
        const uint64_t* mul = P2V_FT_TAB(RYU_POW5_INV)[q];
        vr = ryu_mulshift(4 * m2, mul, i);
        vp = ryu_mulshift(4 * m2 + 2, mul, i);
        vm = ryu_mulshift(4 * m2 - 1 - mm_shift, mul, i);
        if (q <= 21) {
            const uint32_t mod5 = (uint32_t)(mv - 5 * (mv / 5));
            if (mod5 == 0) vr_tz = ryu_pow5factor(mv) >= q;
            else if (accept) vm_tz = ryu_pow5factor(mv - 1 - mm_shift) >= q;
            else vp -= (ryu_pow5factor(mv + 2) >= q) ? 1u : 0u;
        }
    } else {
        const uint32_t q = ryu_log10pow5(-e2) - (-e2 > 1 ? 1u : 0u);
        e10 = (int32_t)q + e2;
        const int32_t i = -e2 - (int32_t)q;
        const int32_t k = (int32_t)ryu_pow5bits(i) - 125;
        const int32_t j = (int32_t)q - k;
        const uint64_t* mul = P2V_FT_TAB(RYU_POW5)[i];
        vr = ryu_mulshift(4 * m2, mul, j);
        vp = ryu_mulshift(4 * m2 + 2, mul, j);
        vm = ryu_mulshift(4 * m2 - 1 - mm_shift, mul, j);
        if (q <= 1) {
            vr_tz = true;
            if (accept) vm_tz = mm_shift == 1;
            else --vp;
        } else if (q < 63) {
            vr_tz = (mv & ((1ull << q) - 1)) == 0;
        }
    }
    int32_t removed = 0;
    uint32_t last = 0;
    uint64_t out;
    if (vm_tz || vr_tz) {
        for (;;) {
            const uint64_t vp10 = vp / 10, vm10 = vm / 10;
            if (vp10 <= vm10) break;
            const uint32_t vm_mod = (uint32_t)(vm - 10 * vm10);
            const uint64_t vr10 = vr / 10;
            const uint32_t vr_mod = (uint32_t)(vr - 10 * vr10);
            vm_tz &= vm_mod == 0;
            vr_tz &= last == 0;
            last = vr_mod;
            vr = vr10; vp = vp10; vm = vm10;
            ++removed;
        }
        if (vm_tz) {
            for (;;) {
                const uint64_t vm10 = vm / 10;
                const uint32_t vm_mod = (uint32_t)(vm - 10 * vm10);
                if (vm_mod != 0) break;
                const uint64_t vp10 = vp / 10, vr10 = vr / 10;
                const uint32_t vr_mod = (uint32_t)(vr - 10 * vr10);
                vr_tz &= last == 0;
                last = vr_mod;
                vr = vr10; vp = vp10; vm = vm10;
                ++removed;
            }
        }
        if (vr_tz && last == 5 && vr % 2 == 0) last = 4;          // exactly half: round to even
        out = vr + (((vr == vm && (!accept || !vm_tz)) || last >= 5) ? 1u : 0u);
    } else {
        bool round_up = false;
        for (;;) {
            const uint64_t vp10 = vp / 10, vm10 = vm / 10;
            if (vp10 <= vm10) break;
            const uint64_t vr10 = vr / 10;
            round_up = (uint32_t)(vr - 10 * vr10) >= 5;
            vr = vr10; vp = vp10; vm = vm10;
            ++removed;
        }
        out = vr + ((vr == vm || round_up) ? 1u : 0u);
    }
    Decimal64 d;
    d.mantissa = out;
    d.exponent = e10 + removed;
    return d;
}

P2V_HD int dec_len17(uint64_t v) {
    int n = 1;
    while (v >= 10) { v /= 10; ++n; }
    return n;
}
P2V_HD int put_uint(char* p, uint64_t v, int ndig) {       // ndig digits of v, most significant first
    for (int i = ndig - 1; i >= 0; --i) { p[i] = (char)('0' + (int)(v % 10)); v /= 10; }
    return ndig;
}

// Writes the string (no terminator) and returns its length (<= 24).
P2V_HD int f64_to_shortest(double x, char* buf) {
    const uint64_t bits = f64_bits(x);
    const bool neg = (bits >> 63) != 0;
    const uint64_t man = bits & ((1ull << 52) - 1);
    const uint32_t ex = (uint32_t)((bits >> 52) & 0x7FFu);
    int n = 0;
    if (ex == 0x7FFu) {
        if (man) { buf[0] = 'N'; buf[1] = 'a'; buf[2] = 'N'; return 3; }
        if (neg) buf[n++] = '-';
        buf[n++] = 'i'; buf[n++] = 'n'; buf[n++] = 'f';
        return n;
    }
    if (neg) buf[n++] = '-';
    if (ex == 0 && man == 0) { buf[n++] = '0'; buf[n++] = '.'; buf[n++] = '0'; return n; }
    const Decimal64 d = ryu_d2d(man, ex);
    const int length = dec_len17(d.mantissa);
    const int k = d.exponent;
    const int kk = length + k;                        // 10^(kk-1) <= |x| < 10^kk
    if (0 <= k && kk <= 16) {                         // 1234e7 -> 12340000000.0
        n += put_uint(buf + n, d.mantissa, length);
        for (int i = length; i < kk; ++i) buf[n++] = '0';
        buf[n++] = '.'; buf[n++] = '0';
    } else if (0 < kk && kk <= 16) {                  // 1234e-2 -> 12.34
        char tmp[20];
        put_uint(tmp, d.mantissa, length);
        for (int i = 0; i < kk; ++i) buf[n++] = tmp[i];
        buf[n++] = '.';
        for (int i = kk; i < length; ++i) buf[n++] = tmp[i];
    } else if (-5 < kk && kk <= 0) {                  // 1234e-6 -> 0.001234
        buf[n++] = '0'; buf[n++] = '.';
        for (int i = kk; i < 0; ++i) buf[n++] = '0';
        n += put_uint(buf + n, d.mantissa, length);
    } else {                                          // 1e30, 1.234e33
        char tmp[20];
        put_uint(tmp, d.mantissa, length);
        buf[n++] = tmp[0];
        if (length > 1) {
            buf[n++] = '.';
            for (int i = 1; i < length; ++i) buf[n++] = tmp[i];
        }
        buf[n++] = 'e';
        int e = kk - 1;
        if (e < 0) { buf[n++] = '-'; e = -e; }
        n += put_uint(buf + n, (uint64_t)e, e >= 100 ? 3 : (e >= 10 ? 2 : 1));
    }
    return n;
}

// ------------------------------------------------------------------------------------------ Eisel-Lemire
// Correctly rounded double of w * 10^q (w != 0).
P2V_HD double el_compute(int64_t q, uint64_t w) {
    if (q < -342) return 0.0;
    if (q > 308) return bits_f64(0x7FF0000000000000ull);
    const int lz = clz64(w);
    w <<= lz;
    const uint64_t* p5 = P2V_FT_TAB(EL_POW5)[q + 342];
    uint64_t hi = umulh(w, p5[0]), lo = w * p5[0];
    if ((hi & 0x1FFull) == 0x1FFull) {               // not enough bits: add the low word of the power
        const uint64_t second_hi = umulh(w, p5[1]);
        lo += second_hi;
        if (second_hi > lo) ++hi;
    }
    const int upperbit = (int)(hi >> 63);
    const int shift = upperbit + 64 - 52 - 3;
    uint64_t mantissa = hi >> shift;
    int64_t power2 = ((217706 * q) >> 16) + 63 + upperbit - lz + 1023;
    if (power2 <= 0) {                                // subnormal (or zero)
        if (-power2 + 1 >= 64) return 0.0;
        mantissa >>= -power2 + 1;
        mantissa += (mantissa & 1);
        mantissa >>= 1;
        power2 = (mantissa < (1ull << 52)) ? 0 : 1;
        return bits_f64(((uint64_t)power2 << 52) | (mantissa & ((1ull << 52) - 1)));
    }
    if (lo <= 1 && q >= -4 && q <= 23 && (mantissa & 3) == 1) {
        if ((mantissa << shift) == hi) mantissa &= ~1ull;            // exactly half way: to even
    }
    mantissa += (mantissa & 1);
    mantissa >>= 1;
    if (mantissa >= (2ull << 52)) { mantissa = 1ull << 52; ++power2; }
    mantissa &= ~(1ull << 52);
    if (power2 >= 0x7FF) return bits_f64(0x7FF0000000000000ull);
    return bits_f64(((uint64_t)power2 << 52) | mantissa);
}

enum { PARSE_F64_OK = 0, PARSE_F64_INVALID = 1, PARSE_F64_UNDECIDED = 2 };

P2V_HD char ft_lower(char c) { return (c >= 'A' && c <= 'Z') ? (char)(c - 'A' + 'a') : c; }
P2V_HD bool ft_word(const char* s, int len, const char* w, int wl) {
    if (len != wl) return false;
    for (int i = 0; i < wl; ++i) if (ft_lower(s[i]) != w[i]) return false;
    return true;
}

P2V_HD int parse_f64(const char* s, int len, double* out) {
    int i = 0;
    bool neg = false;
    if (i < len && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; ++i; }
    if (i >= len) return PARSE_F64_INVALID;
    if (ft_word(s + i, len - i, "inf", 3) || ft_word(s + i, len - i, "infinity", 8)) {
        *out = bits_f64((neg ? 0x8000000000000000ull : 0ull) | 0x7FF0000000000000ull);
        return PARSE_F64_OK;
    }
    if (ft_word(s + i, len - i, "nan", 3)) { *out = bits_f64(0x7FF8000000000000ull); return PARSE_F64_OK; }
    uint64_t w = 0;
    int nd = 0;                 // significant digits taken into w (<= 19)
    int64_t exp10 = 0;          // decimal exponent still to apply to w
    bool any_digit = false, dropped_nonzero = false;
    for (; i < len && s[i] >= '0' && s[i] <= '9'; ++i) {
        any_digit = true;
        const int dgt = s[i] - '0';
        if (nd == 0 && dgt == 0) continue;            // leading zeros
        if (nd < 19) { w = w * 10 + (uint64_t)dgt; ++nd; }
        else { ++exp10; dropped_nonzero |= dgt != 0; }
    }
    if (i < len && s[i] == '.') {
        ++i;
        for (; i < len && s[i] >= '0' && s[i] <= '9'; ++i) {
            any_digit = true;
            const int dgt = s[i] - '0';
            if (nd == 0 && dgt == 0) { --exp10; continue; }
            if (nd < 19) { w = w * 10 + (uint64_t)dgt; ++nd; --exp10; }
            else dropped_nonzero |= dgt != 0;
        }
    }
    if (!any_digit) return PARSE_F64_INVALID;
    if (i < len && (s[i] == 'e' || s[i] == 'E')) {
        ++i;
        bool eneg = false;
        if (i < len && (s[i] == '+' || s[i] == '-')) { eneg = s[i] == '-'; ++i; }
        if (i >= len) return PARSE_F64_INVALID;
        int64_t e = 0;
        bool edig = false;
        for (; i < len && s[i] >= '0' && s[i] <= '9'; ++i) {
            edig = true;
            if (e < 100000) e = e * 10 + (s[i] - '0');
        }
        if (!edig) return PARSE_F64_INVALID;
        exp10 += eneg ? -e : e;
    }
    if (i != len) return PARSE_F64_INVALID;
    double v;
    if (w == 0) v = 0.0;
    else {
        v = el_compute(exp10, w);
        if (dropped_nonzero) {                        // more than 19 digits: the cut and the cut + 1 must agree
            const double v1 = el_compute(exp10, w + 1);
            if (f64_bits(v) != f64_bits(v1)) return PARSE_F64_UNDECIDED;
        }
    }
    *out = neg ? -v : v;
    return PARSE_F64_OK;
}

// ------------------------------------------------------------------------- exact decision for long literals
// parse_f64 leaves a literal of more than 19 significant digits undecided when its 19-digit cut and the cut + 1
// round to different doubles (about one such literal in a thousand): the literal then lies within 1e-19 (relative)
// of the midpoint of two adjacent doubles v < v1.  parse_f64_exact decides it exactly: all significant digits D
// (up to FT_MAX_DIGITS; the midpoint of two doubles has at most 767) and the midpoint (2m + 1) 2^(e-1) are
// brought to integers -- multiply by 5^|E10| on one side, shift by the powers of two -- and compared as big
// numbers in caller-supplied scratch (2 x FT_BIG_LIMBS words; the callers keep one per warp or CTA, the slow
// path is run by one thread at a time).  Ties go to the even mantissa, like Rust's str::parse::<f64>.
constexpr int FT_MAX_DIGITS = 800;
constexpr int FT_BIG_LIMBS = 96;              // 3072 bits: both sides stay below 2750 bits

P2V_HD bool bn_mul_add(uint32_t* a, int& n, uint32_t mul, uint32_t add) {
    uint64_t carry = add;
    for (int i = 0; i < n; ++i) {
        const uint64_t t = (uint64_t)a[i] * mul + carry;
        a[i] = (uint32_t)t; carry = t >> 32;
    }
    if (carry) { if (n >= FT_BIG_LIMBS) return false; a[n++] = (uint32_t)carry; }
    return true;
}
P2V_HD bool bn_mul_pow5(uint32_t* a, int& n, int64_t k) {
    for (; k >= 13; k -= 13) if (!bn_mul_add(a, n, 1220703125u, 0u)) return false;      // 5^13
    uint32_t r = 1;
    for (; k > 0; --k) r *= 5u;
    return r == 1u || bn_mul_add(a, n, r, 0u);
}
P2V_HD bool bn_shl(uint32_t* a, int& n, int64_t bits) {
    if (n == 0 || bits <= 0) return true;
    const int words = (int)(bits >> 5), sh = (int)(bits & 31);
    if (n + words + 1 > FT_BIG_LIMBS) return false;
    uint32_t top = 0;
    if (sh) {
        for (int i = 0; i < n; ++i) { const uint32_t x = a[i]; a[i] = (x << sh) | top; top = x >> (32 - sh); }
        if (top) a[n++] = top;
    }
    if (words) {
        for (int i = n - 1; i >= 0; --i) a[i + words] = a[i];
        for (int i = 0; i < words; ++i) a[i] = 0;
        n += words;
    }
    return true;
}
P2V_HD int bn_cmp(const uint32_t* a, int na, const uint32_t* b, int nb) {
    if (na != nb) return na > nb ? 1 : -1;
    for (int i = na - 1; i >= 0; --i) if (a[i] != b[i]) return a[i] > b[i] ? 1 : -1;
    return 0;
}

// s[0 .. len): a literal parse_f64 returned PARSE_F64_UNDECIDED for (any valid literal works).  A, Bn: FT_BIG_LIMBS words each.
P2V_HD int parse_f64_exact(const char* s, int len, uint32_t* A, uint32_t* Bn, double* out) {
    int i = 0;
    bool neg = false;
    if (i < len && (s[i] == '+' || s[i] == '-')) { neg = s[i] == '-'; ++i; }
    uint64_t w = 0;
    int nd = 0, nbig = 0, na = 0;       // digits in w (<= 19), digits in A (<= FT_MAX_DIGITS), limbs of A
    int64_t exp19 = 0, expbig = 0;      // decimal exponents that go with w and with A
    bool any_digit = false, sticky = false;
    auto take = [&](int dgt, bool frac) -> bool {
        if (nd == 0 && dgt == 0) { if (frac) { --exp19; --expbig; } return true; }          // leading zeros
        if (nd < 19) { w = w * 10 + (uint64_t)dgt; ++nd; if (frac) --exp19; } else if (!frac) ++exp19;
        if (nbig < FT_MAX_DIGITS) { if (!bn_mul_add(A, na, 10u, (uint32_t)dgt)) return false; ++nbig; if (frac) --expbig; }
        else { sticky |= dgt != 0; if (!frac) ++expbig; }
        return true;
    };
    for (; i < len && s[i] >= '0' && s[i] <= '9'; ++i) { any_digit = true; if (!take(s[i] - '0', false)) return PARSE_F64_UNDECIDED; }
    if (i < len && s[i] == '.') {
        ++i;
        for (; i < len && s[i] >= '0' && s[i] <= '9'; ++i) { any_digit = true; if (!take(s[i] - '0', true)) return PARSE_F64_UNDECIDED; }
    }
    if (!any_digit) return PARSE_F64_INVALID;
    if (i < len && (s[i] == 'e' || s[i] == 'E')) {
        ++i;
        bool eneg = false;
        if (i < len && (s[i] == '+' || s[i] == '-')) { eneg = s[i] == '-'; ++i; }
        if (i >= len) return PARSE_F64_INVALID;
        int64_t e = 0;
        bool edig = false;
        for (; i < len && s[i] >= '0' && s[i] <= '9'; ++i) { edig = true; if (e < 100000) e = e * 10 + (s[i] - '0'); }
        if (!edig) return PARSE_F64_INVALID;
        exp19 += eneg ? -e : e; expbig += eneg ? -e : e;
    }
    if (i != len) return PARSE_F64_INVALID;
    if (w == 0) { *out = neg ? -0.0 : 0.0; return PARSE_F64_OK; }
    const double v = el_compute(exp19, w), v1 = el_compute(exp19, w + 1);
    double r = v;
    if (f64_bits(v) != f64_bits(v1)) {
        // v = m 2^e (m < 2^53), its successor v1 = (m + 1) 2^e: the midpoint is (2m + 1) 2^(e - 1)
        const uint64_t bits = f64_bits(v);
        const int ef = (int)((bits >> 52) & 0x7FF);
        const uint64_t m = (bits & ((1ull << 52) - 1)) | (ef ? (1ull << 52) : 0ull);
        const int64_t e2 = (ef ? ef - 1075 : -1074) - 1;
        const uint64_t mid = 2 * m + 1;
        int nb = 0;
        Bn[nb++] = (uint32_t)mid;
        if (mid >> 32) Bn[nb++] = (uint32_t)(mid >> 32);
        // A 10^expbig  against  Bn 2^e2
        bool fits = true;
        if (expbig >= 0) {
            fits = bn_mul_pow5(A, na, expbig);
            const int64_t sh = e2 - expbig;                       // A 5^E  against  Bn 2^(e2 - E)
            fits = fits && (sh >= 0 ? bn_shl(Bn, nb, sh) : bn_shl(A, na, -sh));
        } else {
            const int64_t F = -expbig, t = e2 + F;                // A  against  Bn 5^F 2^(e2 + F)
            fits = bn_mul_pow5(Bn, nb, F);
            fits = fits && (t >= 0 ? bn_shl(Bn, nb, t) : bn_shl(A, na, -t));
        }
        if (!fits) return PARSE_F64_UNDECIDED;
        const int c = bn_cmp(A, na, Bn, nb);
        if (c > 0 || (c == 0 && (sticky || (m & 1)))) r = v1;
    }
    *out = neg ? -r : r;
    return PARSE_F64_OK;
}

}  // namespace p2v_ft
