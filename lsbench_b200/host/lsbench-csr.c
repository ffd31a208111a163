/*
 * lsbench-csr.c -- matrix ingest for the harness: COO text -> host CSR.
 *
 * Same observable behaviour as the reference reader (src/lsbench-csr.c:29-92):
 * header "nnz base\n" with base in {0,1} and nnz > 0 (:37-43); nnz records
 * "row col val\n", the newline mandatory on every record (:49-53); entries
 * ordered by (row, col) (:54), duplicates summed (:57-63); nrows = number of
 * distinct row ids, i.e. absent rows are compressed away (:66-70); offsets
 * 0-based, columns keep the file's base (:79-86).  tests/ checks the result
 * bit for bit against the reference's own reader.
 *
 * Own implementation: one read of the whole file, strtoul/strtod tokens, and
 * a stable bottom-up merge sort that is skipped when the file is already
 * ordered (every matrix the reference ships is).  Pseudo file names describe
 * synthetic operators that only exist on the device (see lsbench.h).
 */
#define _GNU_SOURCE
#include "lsbench-impl.h"
#include <errno.h>
#include <stdint.h>
#include <string.h>

struct entry {
  uint64_t key; /* row << 32 | col */
  double val;
};

static void merge_sort(struct entry *a, size_t n) {
  struct entry *tmp = tcalloc(struct entry, n ? n : 1), *src = a, *dst = tmp;
  if (tmp == NULL)
    err(EXIT_FAILURE, "Unable to allocate sort buffer");
  for (size_t w = 1; w < n; w *= 2) {
    for (size_t lo = 0; lo < n; lo += 2 * w) {
      size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      size_t i = lo, j = mid, k = lo;
      while (i < mid && j < hi)
        dst[k++] = src[j].key < src[i].key ? src[j++] : src[i++];
      while (i < mid)
        dst[k++] = src[i++];
      while (j < hi)
        dst[k++] = src[j++];
    }
    struct entry *t = src;
    src = dst, dst = t;
  }
  if (src != a)
    memcpy(a, src, n * sizeof(struct entry));
  tfree(tmp);
}

static struct csr *synthetic(const char *name) {
  /* poisson7:N | poisson27:N | powerlaw:n[:seed] */
  static const struct {
    const char *prefix;
    int kind;
  } kinds[] = {{"poisson7:", 1}, {"poisson27:", 2}, {"powerlaw:", 3}};
  for (unsigned k = 0; k < 3; k++) {
    size_t len = strlen(kinds[k].prefix);
    if (strncmp(name, kinds[k].prefix, len) != 0)
      continue;
    char *end;
    unsigned long long size = strtoull(name + len, &end, 10), seed = 1;
    if (end == name + len || size == 0)
      errx(EXIT_FAILURE, "Bad synthetic matrix name \"%s\".", name);
    if (*end == ':')
      seed = strtoull(end + 1, NULL, 10);
    unsigned long long n = kinds[k].kind == 3 ? size : size * size * size;
    if (n == 0 || n >= 0xffffffffull)
      errx(EXIT_FAILURE, "Synthetic matrix \"%s\" has too many rows.", name);
    struct csr *A = tcalloc(struct csr, 1);
    A->nrows = (unsigned)n, A->base = 0;
    A->gen_kind = kinds[k].kind, A->gen_size = size, A->gen_seed = seed;
    return A;
  }
  return NULL;
}

struct csr *lsbench_matrix_read(const char *fname) {
  struct csr *S = synthetic(fname);
  if (S)
    return S;

  FILE *fp = fopen(fname, "rb");
  if (!fp)
    err(EXIT_FAILURE, "Unable to open file \"%s\" for reading", fname);
  if (fseek(fp, 0, SEEK_END) != 0)
    err(EXIT_FAILURE, "Unable to seek in \"%s\"", fname);
  long size = ftell(fp);
  rewind(fp);
  char *text = (char *)malloc((size_t)size + 1);
  if (!text || fread(text, 1, (size_t)size, fp) != (size_t)size)
    err(EXIT_FAILURE, "Unable to read \"%s\"", fname);
  text[size] = '\0';
  fclose(fp);

  char *p = text, *q;
  errno = 0;
  unsigned long nnz_l = strtoul(p, &q, 10);
  int ok = q != p;
  p = q;
  unsigned long base_l = strtoul(p, &q, 10);
  ok = ok && q != p && *q == '\n' && !errno;
  if (!ok)
    errx(EXIT_FAILURE, "Unable to read meta information about the matrix.");
  if (base_l > 1)
    errx(EXIT_FAILURE, "Base should be either 0 or 1, got: %lu.", base_l);
  if (nnz_l == 0)
    errx(EXIT_FAILURE, "Number of nnz values in the file are zero.");
  p = q + 1;
  size_t nnz = nnz_l;

  struct entry *a = tcalloc(struct entry, nnz);
  if (a == NULL)
    err(EXIT_FAILURE, "Unable to allocate memories for %zu COO entries.", nnz);
  int sorted = 1;
  for (size_t i = 0; i < nnz; i++) {
    unsigned long r = strtoul(p, &q, 10);
    int good = q != p;
    p = q;
    unsigned long c = strtoul(p, &q, 10);
    good = good && q != p;
    p = q;
    double v = strtod(p, &q);
    if (!good || q == p || *q != '\n')
      errx(EXIT_FAILURE, "Unable to read matrix entries.");
    p = q + 1;
    a[i].key = ((uint64_t)(unsigned)r << 32) | (unsigned)c, a[i].val = v;
    if (i && a[i].key < a[i - 1].key)
      sorted = 0;
  }
  free(text);
  if (!sorted)
    merge_sort(a, nnz);

  /* fold duplicates in place, left to right */
  size_t m = 0;
  for (size_t s = 0; s < nnz; m++) {
    size_t e = s + 1;
    a[m] = a[s];
    for (; e < nnz && a[e].key == a[m].key; e++)
      a[m].val += a[e].val;
    s = e;
  }

  unsigned nrows = 1;
  for (size_t i = 1; i < m; i++)
    nrows += (a[i].key >> 32) != (a[i - 1].key >> 32);

  struct csr *A = tcalloc(struct csr, 1);
  A->nrows = nrows, A->base = (unsigned)base_l;
  A->offs = tcalloc(unsigned, (size_t)nrows + 1);
  A->cols = tcalloc(unsigned, m);
  A->vals = tcalloc(double, m);
  if (!A->offs || !A->cols || !A->vals)
    err(EXIT_FAILURE, "Unable to allocate the CSR arrays");
  unsigned row = 0;
  for (size_t i = 0; i < m; i++) {
    if (i && (a[i].key >> 32) != (a[i - 1].key >> 32))
      A->offs[++row] = (unsigned)i;
    A->cols[i] = (unsigned)(a[i].key & 0xffffffffu);
    A->vals[i] = a[i].val;
  }
  A->offs[nrows] = (unsigned)m;
  tfree(a);
  return A;
}

void lsbench_matrix_print(const struct csr *A) {
  if (!A || !A->offs)
    return;
  for (unsigned i = 0; i < A->nrows; i++)
    for (unsigned k = A->offs[i]; k < A->offs[i + 1]; k++)
      printf("%u %u %lf\n", i + A->base, A->cols[k], A->vals[k]);
}

void lsbench_matrix_free(struct csr *A) {
  if (!A)
    return;
  tfree(A->offs), tfree(A->cols), tfree(A->vals);
  tfree(A);
}

unsigned lsbench_matrix_rows(const struct csr *A) { return A->nrows; }

unsigned long long lsbench_matrix_nnz(const struct csr *A) {
  return A->offs ? A->offs[A->nrows] : A->gen_nnz;
}
