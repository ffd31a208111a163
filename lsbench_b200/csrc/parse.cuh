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
//              are entirely on this path.  Anything else -- more digits, a
//              larger exponent, inf / nan / hex floats, stray characters, a
//              missing newline -- is not guessed at.
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
  if (w > (1ull << 53) || q < -22 || q > 22)
    return false;
  const double m = (double)w;  // exact
  const double r = q < 0 ? m / b2_pow10_exact(-q) : m * b2_pow10_exact(q);
  out = neg ? -r : r;
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
