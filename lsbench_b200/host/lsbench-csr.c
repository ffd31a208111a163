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
 * SURVEY 8(f) row 1 extensions, both opt-in by environment variable: a binary
 * cache of the parsed CSR (LSBENCH_MATRIX_CACHE) and a device-side sort /
 * fold / row-compress (LSBENCH_B200_INGEST, b200 builds only).
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

/* ---- binary cache -------------------------------------
 * With LSBENCH_MATRIX_CACHE set, the CSR a text file parses to is kept beside
 * it as "<file>.lsbcsr" and loaded from there the next time, skipping the
 * tokeniser and the sort; the cache is ignored when older than the text file.
 * Layout: 8-byte magic, u32 nrows, u32 base, u64 nnz, then offs[nrows+1] u32,
 * cols[nnz] u32 (base kept), vals[nnz] f64 -- the members of `struct csr`. */
#include <sys/stat.h>
static const char cache_magic[8] = {'L', 'S', 'B', 'C', 'S', 'R', '1', 0};

static char *cache_name(const char *fname) {
  size_t n = strlen(fname);
  char *c = (char *)malloc(n + 8);
  if (c) {
    memcpy(c, fname, n);
    memcpy(c + n, ".lsbcsr", 8);
  }
  return c;
}

static struct csr *cache_load(const char *fname) {
  char *cn = cache_name(fname);
  struct stat st_txt, st_bin;
  struct csr *A = NULL;
  FILE *fp = NULL;
  if (!cn || stat(fname, &st_txt) != 0 || stat(cn, &st_bin) != 0 ||
      st_bin.st_mtime < st_txt.st_mtime || !(fp = fopen(cn, "rb")))
    goto out;
  char magic[8];
  uint32_t hdr[2];
  uint64_t nnz;
  if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, cache_magic, 8) != 0 ||
      fread(hdr, 4, 2, fp) != 2 || fread(&nnz, 8, 1, fp) != 1 || hdr[0] == 0 ||
      hdr[1] > 1 || nnz == 0 || nnz > 0xffffffffull ||
      (uint64_t)st_bin.st_size != 24 + 4 * ((uint64_t)hdr[0] + 1) + 12 * nnz)
    goto out;
  A = tcalloc(struct csr, 1);
  A->nrows = hdr[0], A->base = hdr[1];
  A->offs = tcalloc(unsigned, (size_t)hdr[0] + 1);
  A->cols = tcalloc(unsigned, nnz);
  A->vals = tcalloc(double, nnz);
  if (!A->offs || !A->cols || !A->vals ||
      fread(A->offs, 4, (size_t)hdr[0] + 1, fp) != (size_t)hdr[0] + 1 ||
      fread(A->cols, 4, nnz, fp) != nnz || fread(A->vals, 8, nnz, fp) != nnz ||
      A->offs[0] != 0 || A->offs[hdr[0]] != nnz) {
    lsbench_matrix_free(A);
    A = NULL;
  }
out:
  if (fp)
    fclose(fp);
  free(cn);
  return A;
}

static void cache_store(const char *fname, const struct csr *A) {
  char *cn = cache_name(fname);
  FILE *fp = cn ? fopen(cn, "wb") : NULL;
  if (fp) {
    uint32_t hdr[2] = {A->nrows, A->base};
    uint64_t nnz = A->offs[A->nrows];
    int ok = fwrite(cache_magic, 1, 8, fp) == 8 && fwrite(hdr, 4, 2, fp) == 2 &&
             fwrite(&nnz, 8, 1, fp) == 1 &&
             fwrite(A->offs, 4, (size_t)A->nrows + 1, fp) == (size_t)A->nrows + 1 &&
             fwrite(A->cols, 4, nnz, fp) == nnz && fwrite(A->vals, 8, nnz, fp) == nnz;
    if (fclose(fp) != 0 || !ok)
      remove(cn); /* a cache that cannot be written is not an error */
  }
  free(cn);
}

#if defined(LSBENCH_B200)
#include "b200.h"
/* LSBENCH_B200_INGEST=1: the sort / fold / row-compress / fill body of the
 * reader (src/lsbench-csr.c:54-86) runs on the GPU (b200_coo_to_csr).  No
 * fallback: a failure here is fatal, like every device error of the backend. */
static struct csr *device_ingest(size_t nnz, unsigned base, const uint32_t *rows,
                                 const uint32_t *cols, const double *vals) {
  const char *v = getenv("LSBENCH_B200_DEVICE");
  b200_ctx *ctx = NULL;
  if (b200_ctx_create(v ? atoi(v) : 0, &ctx) != B200_OK)
    errx(EXIT_FAILURE, "b200 ingest: %s", b200_last_error());
  struct csr *A = tcalloc(struct csr, 1);
  A->base = base;
  A->offs = tcalloc(unsigned, nnz + 1);
  A->cols = tcalloc(unsigned, nnz);
  A->vals = tcalloc(double, nnz);
  if (!A->offs || !A->cols || !A->vals)
    err(EXIT_FAILURE, "Unable to allocate the CSR arrays");
  uint32_t nrows = 0;
  uint64_t m = 0;
  if (b200_coo_to_csr(ctx, nnz, rows, cols, vals, &nrows, &m, A->offs, A->cols,
                      A->vals) != B200_OK)
    errx(EXIT_FAILURE, "b200 ingest: %s", b200_last_error());
  b200_ctx_destroy(ctx);
  A->nrows = nrows;
  /* give back what the duplicates and absent rows did not need */
  unsigned *o = (unsigned *)realloc(A->offs, ((size_t)nrows + 1) * sizeof(unsigned));
  unsigned *c = (unsigned *)realloc(A->cols, (m ? m : 1) * sizeof(unsigned));
  double *d = (double *)realloc(A->vals, (m ? m : 1) * sizeof(double));
  A->offs = o ? o : A->offs, A->cols = c ? c : A->cols, A->vals = d ? d : A->vals;
  return A;
}
#endif

#if defined(LSBENCH_B200)
/* Text straight to CSR on the device (b200_text_to_csr).  NULL when the body
 * is not one strict record per line (the caller tokenises it instead); any
 * other failure is fatal. */
static struct csr *device_text_ingest(const char *body, size_t len, size_t nnz,
                                      unsigned base) {
  const char *v = getenv("LSBENCH_B200_DEVICE");
  b200_ctx *ctx = NULL;
  if (b200_ctx_create(v ? atoi(v) : 0, &ctx) != B200_OK)
    errx(EXIT_FAILURE, "b200 ingest: %s", b200_last_error());
  struct csr *A = tcalloc(struct csr, 1);
  A->base = base;
  A->offs = tcalloc(unsigned, nnz + 1);
  A->cols = tcalloc(unsigned, nnz);
  A->vals = tcalloc(double, nnz);
  if (!A->offs || !A->cols || !A->vals)
    err(EXIT_FAILURE, "Unable to allocate the CSR arrays");
  uint32_t nrows = 0;
  uint64_t m = 0, nhost = 0;
  int rc = b200_text_to_csr(ctx, body, len, nnz, &nrows, &m, A->offs, A->cols,
                            A->vals, &nhost);
  b200_ctx_destroy(ctx);
  if (rc == B200_EINVAL) { /* not strict: let the tokeniser judge the file */
    lsbench_matrix_free(A);
    return NULL;
  }
  if (rc != B200_OK)
    errx(EXIT_FAILURE, "b200 ingest: %s", b200_last_error());
  A->nrows = nrows;
  unsigned *o = (unsigned *)realloc(A->offs, ((size_t)nrows + 1) * sizeof(unsigned));
  unsigned *c = (unsigned *)realloc(A->cols, (m ? m : 1) * sizeof(unsigned));
  double *d = (double *)realloc(A->vals, (m ? m : 1) * sizeof(double));
  A->offs = o ? o : A->offs, A->cols = c ? c : A->cols, A->vals = d ? d : A->vals;
  return A;
}
#endif

struct csr *lsbench_matrix_read(const char *fname) {
  struct csr *S = synthetic(fname);
  if (S)
    return S;
  const char *cache_env = getenv("LSBENCH_MATRIX_CACHE");
  const int use_cache = cache_env && *cache_env && strcmp(cache_env, "0") != 0;
  if (use_cache && (S = cache_load(fname)))
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

#if defined(LSBENCH_B200)
  {
    /* LSBENCH_B200_INGEST: lines found and parsed on the GPU when the body is
     * one strict record per line; otherwise the tokeniser below */
    const char *ing = getenv("LSBENCH_B200_INGEST");
    if (ing && *ing && strcmp(ing, "0") != 0) {
      struct csr *D = device_text_ingest(p, (size_t)size - (size_t)(p - text), nnz,
                                         (unsigned)base_l);
      if (D) {
        free(text);
        if (use_cache)
          cache_store(fname, D);
        return D;
      }
    }
  }
#endif

  /* the records in file order (src/lsbench-csr.c:49-53) */
  uint32_t *rows = tcalloc(uint32_t, nnz), *cols = tcalloc(uint32_t, nnz);
  double *vals = tcalloc(double, nnz);
  if (!rows || !cols || !vals)
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
    rows[i] = (uint32_t)r, cols[i] = (uint32_t)c, vals[i] = v;
    if (i && (rows[i] < rows[i - 1] ||
              (rows[i] == rows[i - 1] && cols[i] < cols[i - 1])))
      sorted = 0;
  }
  free(text);

  struct csr *A = NULL;
#if defined(LSBENCH_B200)
  const char *ing = getenv("LSBENCH_B200_INGEST");
  if (ing && *ing && strcmp(ing, "0") != 0)
    A = device_ingest(nnz, (unsigned)base_l, rows, cols, vals);
#endif
  if (A == NULL) {
    struct entry *a = tcalloc(struct entry, nnz);
    if (a == NULL)
      err(EXIT_FAILURE, "Unable to allocate memories for %zu COO entries.", nnz);
    for (size_t i = 0; i < nnz; i++)
      a[i].key = ((uint64_t)rows[i] << 32) | cols[i], a[i].val = vals[i];
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

    A = tcalloc(struct csr, 1);
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
  }
  tfree(rows), tfree(cols), tfree(vals);
  if (use_cache)
    cache_store(fname, A);
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
