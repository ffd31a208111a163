// parse_check.cpp -- host check of lsbench_b200/csrc/parse.cuh against libc:
// whenever b2_parse_record accepts a line, its three fields must equal what
// (unsigned)strtoul / strtod give, bit for bit.  Usage:
//   parse_check <file.txt>...   every record of COO text files
//   parse_check --random N      N generated records in many formats
// Prints "accepted A of T, mismatches M".
#include "parse.cuh"
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

static long total = 0, accepted = 0, mismatches = 0;

static void check_line(const char *p, const char *nl) {
  total++;
  uint32_t r = 0, c = 0;
  double v = 0;
  if (b2_parse_record(p, nl, r, c, v) != B2_PARSE_OK)
    return;
  accepted++;
  char *q;
  unsigned long lr = strtoul(p, &q, 10);
  unsigned long lc = strtoul(q, &q, 10);
  double lv = strtod(q, &q);
  if (q != nl || (uint32_t)lr != r || (uint32_t)lc != c || memcmp(&lv, &v, 8) != 0) {
    mismatches++;
    if (mismatches < 10)
      fprintf(stderr, "MISMATCH on '%.*s': %u %u %a vs %lu %lu %a\n", (int)(nl - p), p, r, c, v,
              lr, lc, lv);
  }
}

int main(int argc, char **argv) {
  if (argc >= 3 && strcmp(argv[1], "--random") == 0) {
    long n = atol(argv[2]);
    std::mt19937_64 g(12345);
    char buf[256];
    for (long i = 0; i < n; i++) {
      uint64_t a = g(), b = g();
      unsigned row = (unsigned)(a % 4000000000ull), col = (unsigned)(b % 4000000000ull);
      double mant = (double)(g() >> 11) / 9007199254740992.0;      // [0, 1)
      int ex = (int)(g() % 40) - 20;
      double v = (g() & 1 ? -1 : 1) * mant * pow(10.0, ex);
      const char *fmts[] = {"%u %u %.15f\n", "%u %u %.10e\n", "%u %u %.17g\n", "%u %u %.6f\n",
                            "%u %u %g\n",    "%u\t%u  %.12E\n", "%u %u %.0f\n", "%u %u %.15g\n",
                            "%u %u %+.8e\n", "%u %u %.3e\n"};
      int len = snprintf(buf, sizeof buf, fmts[g() % 10], row, col, v);
      check_line(buf, buf + len - 1);
    }
    // hand-picked edge cases
    const char *edge[] = {"1 2 0\n", "1 2 -0\n", "1 2 -0.000\n", "1 2 .5\n", "1 2 5.\n", "1 2 1e22\n",
                          "1 2 1e23\n", "1 2 9007199254740992\n", "1 2 9007199254740993\n",
                          "1 2 0.000000000000000000001\n", "1 2 123456789012345678901234\n",
                          "1 2 1e\n", "1 2 1e+\n", "1 2 inf\n", "1 2 nan\n", "1 2 0x1p3\n", "1 2 1.5 \n",
                          " 1 2 1.5\n", "1 2\n", "\n", "4294967295 4294967296 1\n",
                          "1 2 1E-22\n", "1 2 1E-23\n", "007 08 0009.50\n", "1 2 +3.25\n"};
    for (const char *e : edge)
      check_line(e, e + strlen(e) - 1);
  } else {
    for (int i = 1; i < argc; i++) {
      FILE *f = fopen(argv[i], "rb");
      if (!f) {
        perror(argv[i]);
        return 2;
      }
      fseek(f, 0, SEEK_END);
      long sz = ftell(f);
      rewind(f);
      std::vector<char> t(sz + 1);
      if (fread(t.data(), 1, sz, f) != (size_t)sz)
        return 2;
      t[sz] = 0;
      fclose(f);
      const char *p = (const char *)memchr(t.data(), '\n', sz);  // skip the header
      if (!p)
        return 2;
      p++;
      while (p < t.data() + sz) {
        const char *nl = (const char *)memchr(p, '\n', t.data() + sz - p);
        if (!nl)
          break;
        check_line(p, nl);
        p = nl + 1;
      }
    }
  }
  printf("accepted %ld of %ld, mismatches %ld\n", accepted, total, mismatches);
  return mismatches ? 1 : 0;
}
