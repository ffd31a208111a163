// parse.cuh -- one COO record "row col val\n" (src/lsbench-csr.c:49-53:
// fscanf "%u %u %lf\n") from text, usable on the host and on the device.
//
// Strict and exact, or not at all: the function accepts the record only when
// it can produce bit for bit what strtoul / strtod produce; everything else is
// handed back to the caller (B2_PARSE_ASK_HOST), who runs the libc functions
// on that one line.
//
//   integers   up to 19 digits accumulate in 64 bits without overflow and are
//              narrowed to 32 bits exactly like `(unsigned)strtoul(...)`.
//   doubles    [+-]digits[.digits][(e|E)[+-]digits].  With at most 19
//              significant digits the decimal is w x 10^q with w exact in a
//              u64; when w <= 2^53 and |q| <= 22 both w and 10^|q| are exact
//              doubles and ONE correctly rounded IEEE multiply or divide gives
//              the correctly rounded result (Clinger's fast path) -- the same
//              double strtod returns.  The Nek files ("0.217534912180770")
//              are entirely on this path.  Beyond it, for |q| <= 27 (every
//              "%.17g" value between 1e-27 and 1e46), w x 10^q is rounded
//              with EXACT integer arithmetic: a 154-bit product, or a binary
//              long division by 10^|q| < 2^90 with the remainder as sticky
//              bit, then round-half-even to 53 bits -- no tables, no
//              heuristics, so again what a correctly rounding strtod gives.
//              Anything else -- more digits, a larger exponent, inf / nan /
//              hex floats, stray characters, a missing newline -- is not
//              guessed at.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define B2_PARSE_HD __host__ __device__ __forceinline__
#else
#define B2_PARSE_HD static inline
#endif

enum { B2_PARSE_OK = 0, B2_PARSE_ASK_HOST = 1 };

B2_PARSE_HD bool b2_is_digit(char c) { return c >= '0' && c <= '9'; }
B2_PARSE_HD bool b2_is_blank(char c) { return c == ' ' || c == '\t'; }

// unsigned field at p (no sign, no leading blanks); advances p
B2_PARSE_HD bool b2_parse_u32(const char *&p, const char *end, uint32_t &out) {
  if (p >= end || !b2_is_digit(*p))
    return false;
  uint64_t v = 0;
  int nd = 0;
  while (p < end && b2_is_digit(*p)) {
    if (++nd > 19)
      return false;
    v = v * 10u + (uint64_t)(*p - '0');
    p++;
  }
  out = (uint32_t)v;  // (unsigned)strtoul(...): the low 32 bits
  return true;
}

B2_PARSE_HD double b2_pow10_exact(int k) {  // 10^k, 0 <= k <= 22: exact in fp64
  const double t[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,
                        1e8,  1e9,  1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                        1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  return t[k];
}

typedef unsigned __int128 b2_u128;

B2_PARSE_HD int b2_bitlen64(uint64_t v) {
  int n = 0;
  while (v)
    n++, v >>= 1;
  return n;
}
B2_PARSE_HD int b2_bitlen128(b2_u128 v) {
  const uint64_t hi = (uint64_t)(v >> 64);
  return hi ? 64 + b2_bitlen64(hi) : b2_bitlen64((uint64_t)v);
}

// m (up to 128 bits, m != 0) x 2^e2, plus "there were non-zero bits below"
// (sticky): round half to even to 53 bits and assemble the double.  The
// callers keep the result in the normal range.
B2_PARSE_HD double b2_round_pack(b2_u128 m, int e2, bool sticky, bool neg) {
  const int L = b2_bitlen128(m);
  if (L > 53) {
    const int drop = L - 53;
    const b2_u128 one = 1;
    const b2_u128 rest = m & ((one << drop) - 1), half = one << (drop - 1);
    m >>= drop, e2 += drop;
    const bool up = rest > half || (rest == half && (sticky || (m & 1)));
    if (up) {
      m += 1;
      if (m >> 53)
        m >>= 1, e2 += 1;
    }
  } else if (L < 53) {
    m <<= (53 - L), e2 -= (53 - L);  // exact: nothing was dropped
  }
  const uint64_t mant = (uint64_t)m;  // bit 52 set
  const uint64_t bits = ((uint64_t)neg << 63) | ((uint64_t)(e2 + 52 + 1023) << 52) |
                        (mant & ((1ull << 52) - 1));
  double d;
#ifdef __CUDA_ARCH__
  d = __longlong_as_double((long long)bits);
#else
  __builtin_memcpy(&d, &bits, 8);
#endif
  return d;
}

// w x 10^q, 1 <= w < 2^64, |q| <= 27, correctly rounded, by exact integer
// arithmetic (10^27 < 2^90).
B2_PARSE_HD double b2_exact_decimal(uint64_t w, int q, bool neg) {
  b2_u128 p10 = 1;
  for (int i = 0; i < (q < 0 ? -q : q); i++)
    p10 *= 10u;
  if (q >= 0) {
    // 64-bit x 90-bit product, up to 154 bits: mid holds bits 64.., low bits 0..63
    const b2_u128 lo64 = (b2_u128)(uint64_t)p10 * w;
    const b2_u128 mid = (p10 >> 64) * w + (lo64 >> 64);
    const uint64_t low = (uint64_t)lo64;
    if ((mid >> 64) == 0)  // the product fits 128 bits
      return b2_round_pack((mid << 64) | low, 0, false, neg);
    return b2_round_pack(mid, 64, low != 0, neg);  // >= 65 bits kept, low 64 are sticky
  }
  // w / 10^n: binary long division of w x 2^s by p10 with s such that the
  // quotient has at least 55 bits; a non-zero remainder is the sticky bit
  const int bw = b2_bitlen64(w), bd = b2_bitlen128(p10);
  int s = 55 + bd - bw;
  if (s < 0)
    s = 0;
  b2_u128 rem = 0, quo = 0;
  for (int i = bw - 1; i >= -s; i--) {
    rem = (rem << 1) | (i >= 0 ? (b2_u128)((w >> i) & 1u) : (b2_u128)0);
    quo <<= 1;
    if (rem >= p10)
      rem -= p10, quo |= 1u;
  }
  return b2_round_pack(quo, -s, rem != 0, neg);
}

// floating-point field at p; advances p past it
B2_PARSE_HD bool b2_parse_f64(const char *&p, const char *end, double &out) {
  bool neg = false;
  if (p < end && (*p == '-' || *p == '+'))
    neg = *p == '-', p++;
  uint64_t w = 0;
  int sig = 0;         // significant digits taken into w
  int q = 0;           // decimal exponent of w
  bool any = false, seen_nonzero = false;
  while (p < end && b2_is_digit(*p)) {
    any = true;
    if (*p != '0' || seen_nonzero) {
      seen_nonzero = true;
      if (++sig > 19)
        return false;
      w = w * 10u + (uint64_t)(*p - '0');
    }
    p++;
  }
  if (p < end && *p == '.') {
    p++;
    while (p < end && b2_is_digit(*p)) {
      any = true;
      if (*p != '0' || seen_nonzero) {
        seen_nonzero = true;
        if (++sig > 19)
          return false;
        w = w * 10u + (uint64_t)(*p - '0');
      }
      q--;
      p++;
    }
  }
  if (!any)
    return false;
  if (p < end && (*p == 'e' || *p == 'E')) {
    p++;
    bool eneg = false;
    if (p < end && (*p == '-' || *p == '+'))
      eneg = *p == '-', p++;
    if (p >= end || !b2_is_digit(*p))
      return false;  // "1e" / "1e+": strtod would stop before the e; let libc decide
    int e = 0, ed = 0;
    while (p < end && b2_is_digit(*p)) {
      if (++ed > 4)
        return false;
      e = e * 10 + (*p - '0');
      p++;
    }
    q += eneg ? -e : e;
  }
  if (w == 0) {
    out = neg ? -0.0 : 0.0;
    return true;
  }
  if (w <= (1ull << 53) && q >= -22 && q <= 22) {
    const double m = (double)w;  // exact
    const double r = q < 0 ? m / b2_pow10_exact(-q) : m * b2_pow10_exact(q);
    out = neg ? -r : r;
    return true;
  }
  if (q < -27 || q > 27)
    return false;
  out = b2_exact_decimal(w, q, neg);
  return true;
}

// One record occupying exactly [p, nl], nl pointing at its '\n'.
B2_PARSE_HD int b2_parse_record(const char *p, const char *nl, uint32_t &row, uint32_t &col,
                                double &val) {
  if (!b2_parse_u32(p, nl, row) || p >= nl || !b2_is_blank(*p))
    return B2_PARSE_ASK_HOST;
  while (p < nl && b2_is_blank(*p))
    p++;
  if (!b2_parse_u32(p, nl, col) || p >= nl || !b2_is_blank(*p))
    return B2_PARSE_ASK_HOST;
  while (p < nl && b2_is_blank(*p))
    p++;
  if (!b2_parse_f64(p, nl, val) || p != nl)
    return B2_PARSE_ASK_HOST;
  return B2_PARSE_OK;
}
