/*
 * csr_read.c -- restatement of the reference's COO-text -> CSR ingest.
 * TEST INFRASTRUCTURE (see oracle.h).  Follows src/lsbench-csr.c:29-92:
 *
 *   :37-43  header "nnz base", terminated by '\n'; base in {0,1}; nnz > 0
 *   :49-53  nnz records "row col val", each terminated by '\n'
 *   :54     order by (row, col)
 *   :57-63  equal (row, col) records are summed, left to right
 *   :66-70  nrows = number of DISTINCT row ids (absent rows vanish)
 *   :79-86  offs 0-based; cols keep the file's base; row ids are dropped
 *
 * Written independently: whole-file read + strtoul/strtod tokeniser and a
 * stable LSD radix sort instead of fscanf + qsort.  The reference's qsort
 * leaves the order of equal keys unspecified; here duplicates are summed in
 * file order (glibc's qsort is a merge sort for these sizes, so the two agree
 * in practice -- the _ref comparison in tests/ checks it bit for bit).
 */
#include "oracle.h"
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t key; /* row << 32 | col */
  double val;
} rec_t;

static void radix_sort_recs(rec_t *a, rec_t *tmp, size_t n) {
  /* 8 passes of 8 bits, stable. */
  for (int pass = 0; pass < 8; pass++) {
    size_t cnt[257] = {0};
    int sh = pass * 8;
    int trivial = 1;
    uint64_t first = n ? (a[0].key >> sh) & 0xff : 0;
    for (size_t i = 0; i < n; i++) {
      uint64_t d = (a[i].key >> sh) & 0xff;
      cnt[d + 1]++;
      trivial &= (d == first);
    }
    if (trivial)
      continue;
    for (int d = 0; d < 256; d++)
      cnt[d + 1] += cnt[d];
    for (size_t i = 0; i < n; i++)
      tmp[cnt[(a[i].key >> sh) & 0xff]++] = a[i];
    memcpy(a, tmp, n * sizeof(rec_t));
  }
}

static char *slurp(const char *fname, size_t *len) {
  FILE *fp = fopen(fname, "rb");
  if (!fp)
    return NULL;
  fseek(fp, 0, SEEK_END);
  long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  char *buf = (char *)malloc((size_t)sz + 1);
  if (!buf || fread(buf, 1, (size_t)sz, fp) != (size_t)sz) {
    free(buf), fclose(fp);
    return NULL;
  }
  buf[sz] = '\0', *len = (size_t)sz;
  fclose(fp);
  return buf;
}

/* One unsigned field; mirrors "%u": leading white space (incl. newlines)
 * is skipped.  Returns 0 on success. */
static int take_u32(char **p, uint32_t *out) {
  char *end;
  errno = 0;
  unsigned long v = strtoul(*p, &end, 10);
  if (end == *p || errno)
    return 1;
  *out = (uint32_t)v, *p = end;
  return 0;
}

orc_csr *orc_matrix_read(const char *fname) {
  size_t len = 0;
  char *buf = slurp(fname, &len);
  if (!buf)
    return NULL;

  char *p = buf;
  uint32_t nnz, base;
  if (take_u32(&p, &nnz) || take_u32(&p, &base) || *p != '\n')
    goto fail; /* :37-39 */
  p++;
  if (base > 1 || nnz == 0)
    goto fail; /* :40-43 */

  rec_t *a = (rec_t *)malloc((size_t)nnz * sizeof(rec_t));
  rec_t *tmp = (rec_t *)malloc((size_t)nnz * sizeof(rec_t));
  if (!a || !tmp) {
    free(a), free(tmp);
    goto fail;
  }
  for (uint32_t i = 0; i < nnz; i++) {
    uint32_t r, c;
    char *end;
    if (take_u32(&p, &r) || take_u32(&p, &c)) {
      free(a), free(tmp);
      goto fail;
    }
    double v = strtod(p, &end);
    /* :50-52 every record must end in a newline (also the last one). */
    if (end == p || *end != '\n') {
      free(a), free(tmp);
      goto fail;
    }
    p = end + 1;
    a[i].key = ((uint64_t)r << 32) | c, a[i].val = v;
  }
  radix_sort_recs(a, tmp, nnz); /* :54 */
  free(tmp);

  /* :57-63 fold runs of equal keys. */
  size_t m = 0;
  for (size_t s = 0; s < nnz;) {
    rec_t acc = a[s];
    size_t e = s + 1;
    while (e < nnz && a[e].key == acc.key)
      acc.val += a[e++].val;
    a[m++] = acc, s = e;
  }

  /* :66-70 distinct row ids. */
  uint32_t nrows = 1;
  for (size_t i = 1; i < m; i++)
    nrows += (a[i].key >> 32) != (a[i - 1].key >> 32);

  orc_csr *A = (orc_csr *)calloc(1, sizeof(orc_csr));
  A->nrows = nrows, A->base = base;
  A->offs = (uint32_t *)calloc((size_t)nrows + 1, sizeof(uint32_t));
  A->cols = (uint32_t *)calloc(m, sizeof(uint32_t));
  A->vals = (double *)calloc(m, sizeof(double));
  uint32_t row = 0;
  for (size_t i = 0; i < m; i++) {
    if (i && (a[i].key >> 32) != (a[i - 1].key >> 32))
      A->offs[++row] = (uint32_t)i;
    A->cols[i] = (uint32_t)(a[i].key & 0xffffffffu); /* base kept, :79-86 */
    A->vals[i] = a[i].val;
  }
  A->offs[nrows] = (uint32_t)m;
  free(a), free(buf);
  return A;

fail:
  free(buf);
  return NULL;
}

void orc_matrix_free(orc_csr *A) {
  if (!A)
    return;
  free(A->offs), free(A->cols), free(A->vals), free(A);
}
