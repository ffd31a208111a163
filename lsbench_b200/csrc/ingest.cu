// ingest.cu -- SURVEY 8(f) row 1: device-side COO -> CSR, the step right before
// the hot path.
//
// Stands where the body of the reference reader stands once the text has been
// tokenised (src/lsbench-csr.c:54-86): order the COO records by (row, col)
// (:54, qsort there), sum records with equal (row, col) (:57-63), count the
// DISTINCT row ids -- absent rows are compressed away (:66-70) -- and fill the
// CSR with 0-based offsets and columns that keep the file's base (:79-86).
// The reference does this with a 24-byte record qsort on one core and
// `unsigned nnz`; here it is a stable LSD radix sort of 64-bit keys (skipped
// when the file is already ordered, as every matrix the reference ships is),
// two head-flag scans and one fold kernel, all on the device.
//
// Bit-exactness: the sort is stable, so the records of one (row, col) run stay
// in file order and each run is summed left to right by one thread -- the same
// additions in the same order as the reference's fold loop (glibc qsort is a
// merge sort at these sizes, hence stable as well; tests/ compare against the
// reference's own reader).
#include "common.cuh"
#include "parse.cuh"
#include <cub/cub.cuh>
#include <cerrno>
#include <string>
#include <vector>

#define T256 256
static inline unsigned nblk(uint64_t n, unsigned t = T256) {
  return (unsigned)((n + t - 1) / t);
}

__global__ void k_coo_keys(uint64_t nnz, const uint32_t *__restrict__ rows,
                           const uint32_t *__restrict__ cols, uint64_t *keys,
                           uint32_t *idx, unsigned *unsorted,
                           unsigned long long *key_or) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint64_t k = 0;
  if (i < nnz) {
    k = ((uint64_t)rows[i] << 32) | cols[i];
    keys[i] = k, idx[i] = (uint32_t)i;
    if (i + 1 < nnz) {
      uint64_t kn = ((uint64_t)rows[i + 1] << 32) | cols[i + 1];
      if (kn < k)
        *unsorted = 1u;  // benign race: every writer stores 1
    }
  }
  // which key bits are used at all -> radix passes that can be skipped
  uint64_t w = k;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    w |= __shfl_xor_sync(0xffffffffu, w, o);
  if ((threadIdx.x & 31) == 0 && w)
    atomicOr(key_or, (unsigned long long)w);
}

__global__ void k_head_flags(uint64_t n, const uint64_t *__restrict__ keys,
                             int shift, uint32_t *head) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    head[i] = (i == 0 || (keys[i] >> shift) != (keys[i - 1] >> shift)) ? 1u : 0u;
}

// One thread per (row, col) run: sum the run left to right (file order).
__global__ void k_fold_runs(uint64_t nnz, const uint64_t *__restrict__ keys,
                            const uint32_t *__restrict__ idx,
                            const uint32_t *__restrict__ head,
                            const uint32_t *__restrict__ pos,
                            const double *__restrict__ vals, uint64_t *okeys,
                            double *ovals) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nnz || !head[i])
    return;
  const uint64_t k = keys[i];
  double acc = vals[idx[i]];
  for (uint64_t e = i + 1; e < nnz && keys[e] == k; e++)
    acc += vals[idx[e]];  // src/lsbench-csr.c:59-61
  okeys[pos[i]] = k, ovals[pos[i]] = acc;
}

// Entry j of the folded list: column out, and for the first entry of a row the
// offset of the (compressed) row.
__global__ void k_emit_csr(uint64_t m, const uint64_t *__restrict__ okeys,
                           const uint32_t *__restrict__ rhead,
                           const uint32_t *__restrict__ rpos, uint32_t sub,
                           uint32_t *cols, uint64_t *offs) {
  uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j >= m)
    return;
  cols[j] = (uint32_t)(okeys[j] & 0xffffffffu) - sub;
  if (rhead[j])
    offs[rpos[j]] = j;
}

template <typename T>
static int scan_excl(cudaStream_t s, const T *in, T *out, uint64_t n) {
  void *tmp = nullptr;
  size_t bytes = 0;
  CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, s));
  CU_TRY(cudaMalloc(&tmp, bytes ? bytes : 8));
  CU_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, s));
  CU_TRY(cudaStreamSynchronize(s));
  CU_TRY(cudaFree(tmp));
  return B200_OK;
}

struct IngestTmp {
  void *p[24] = {nullptr};
  int n = 0;
  template <typename T> int get(T **out, size_t count) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
    if (e != cudaSuccess) {
      b200_set_error("ingest: cudaMalloc of %zu bytes failed: %s",
                     count * sizeof(T), cudaGetErrorString(e));
      return e == cudaErrorMemoryAllocation ? B200_ENOMEM : B200_ECUDA;
    }
    p[n++] = q, *out = (T *)q;
    return B200_OK;
  }
  ~IngestTmp() {
    for (int i = 0; i < n; i++)
      cudaFree(p[i]);
  }
};

// Host COO (rows / cols as read from the file, i.e. still carrying the base)
// -> plain device CSR.  `sub` is subtracted from the columns (0 keeps the
// file's base, as `struct csr` does; `base` gives the 0-based ids the device
// layout wants).  Row ids are dropped: row r of the result is the r-th
// distinct row id of the input.
// device core: d_rows / d_cols / d_vals are the records in file order
static int coo_dev_to_plain(b200_ctx *c, uint64_t nnz, const uint32_t *d_rows,
                            const uint32_t *d_cols, const double *d_vals, uint32_t sub,
                            PlainCsr *A, uint32_t *was_sorted) {
  if (nnz == 0 || nnz >= 0xffffffffull)
    B_FAIL(B200_ERANGE, "ingest: nnz=%llu outside (0, 2^32)", (unsigned long long)nnz);
  cudaStream_t s = c->stream;
  IngestTmp T;
  uint32_t *idx, *idx2, *head, *pos;
  uint64_t *keys, *keys2;
  unsigned *d_flag;
  unsigned long long *d_or;
  B_TRY(T.get(&keys, nnz));
  B_TRY(T.get(&idx, nnz));
  B_TRY(T.get(&d_flag, 2));
  B_TRY(T.get(&d_or, 1));
  CU_TRY(cudaMemsetAsync(d_flag, 0, 8, s));
  CU_TRY(cudaMemsetAsync(d_or, 0, 8, s));
  k_coo_keys<<<nblk(nnz), T256, 0, s>>>(nnz, d_rows, d_cols, keys, idx, d_flag, d_or);
  CU_TRY(cudaGetLastError());
  unsigned unsorted = 0;
  unsigned long long key_or = 0;
  CU_TRY(cudaMemcpyAsync(&unsorted, d_flag, 4, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaMemcpyAsync(&key_or, d_or, 8, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  if (was_sorted)
    *was_sorted = !unsorted;

  if (unsorted) {  // src/lsbench-csr.c:54
    B_TRY(T.get(&keys2, nnz));
    B_TRY(T.get(&idx2, nnz));
    // radix passes only over the bits some key uses: columns [0, cb), rows [32, 32+rb)
    int cb = 0, rb = 0;
    for (int b = 0; b < 32; b++) {
      if ((key_or >> b) & 1ull) cb = b + 1;
      if ((key_or >> (32 + b)) & 1ull) rb = b + 1;
    }
    void *tmp = nullptr;
    size_t bytes = 0;
    const uint64_t *kin = keys;
    uint64_t *kout = keys2;
    const uint32_t *vin = idx;
    uint32_t *vout = idx2;
    for (int pass = 0; pass < 2; pass++) {
      int lo = pass == 0 ? 0 : 32, hi = pass == 0 ? cb : 32 + rb;
      if (hi <= lo)
        continue;
      bytes = 0;
      CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout,
                                             nnz, lo, hi, s));
      CU_TRY(cudaMalloc(&tmp, bytes ? bytes : 8));
      cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin,
                                                      vout, nnz, lo, hi, s);
      cudaStreamSynchronize(s);
      cudaFree(tmp);
      CU_TRY(e);
      const uint64_t *tk = kin;
      kin = kout, kout = (uint64_t *)tk;
      const uint32_t *tv = vin;
      vin = vout, vout = (uint32_t *)tv;
    }
    if (kin != keys) {  // result landed in the second buffer
      uint64_t *tk = keys;
      keys = keys2, keys2 = tk;
      uint32_t *tv = idx;
      idx = idx2, idx2 = tv;
    }
  }

  // ---- fold equal (row, col) runs (:57-63) ---------------------------------------
  B_TRY(T.get(&head, nnz + 1));
  B_TRY(T.get(&pos, nnz + 1));
  k_head_flags<<<nblk(nnz), T256, 0, s>>>(nnz, keys, 0, head);
  CU_TRY(cudaMemsetAsync(head + nnz, 0, 4, s));
  B_TRY(scan_excl(s, head, pos, nnz + 1));
  uint32_t m32 = 0;
  CU_TRY(cudaMemcpy(&m32, pos + nnz, 4, cudaMemcpyDeviceToHost));
  const uint64_t m = m32;
  uint64_t *okeys;
  B_TRY(T.get(&okeys, m));
  A->nnz = m;
  CU_TRY(cudaMalloc(&A->cols, (m ? m : 1) * 4));
  CU_TRY(cudaMalloc(&A->vals, (m ? m : 1) * 8));
  k_fold_runs<<<nblk(nnz), T256, 0, s>>>(nnz, keys, idx, head, pos, d_vals, okeys, A->vals);
  CU_TRY(cudaGetLastError());

  // ---- distinct row ids -> compressed rows (:66-70), CSR fill (:79-86) -----------
  k_head_flags<<<nblk(m), T256, 0, s>>>(m, okeys, 32, head);
  CU_TRY(cudaMemsetAsync(head + m, 0, 4, s));
  B_TRY(scan_excl(s, head, pos, m + 1));
  uint32_t nrows = 0;
  CU_TRY(cudaMemcpy(&nrows, pos + m, 4, cudaMemcpyDeviceToHost));
  A->n = nrows;
  CU_TRY(cudaMalloc(&A->offs, (nrows + 1ull) * 8));
  k_emit_csr<<<nblk(m), T256, 0, s>>>(m, okeys, head, pos, sub, A->cols, A->offs);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaMemcpyAsync(A->offs + nrows, &A->nnz, 8, cudaMemcpyHostToDevice, s));
  CU_TRY(cudaStreamSynchronize(s));
  return B200_OK;
}

int coo_to_plain(b200_ctx *c, uint64_t nnz, const uint32_t *h_rows,
                 const uint32_t *h_cols, const double *h_vals, uint32_t sub,
                 PlainCsr *A, uint32_t *was_sorted) {
  if (nnz == 0 || nnz >= 0xffffffffull)
    B_FAIL(B200_ERANGE, "ingest: nnz=%llu outside (0, 2^32)", (unsigned long long)nnz);
  cudaStream_t s = c->stream;
  IngestTmp T;
  uint32_t *d_rows, *d_cols;
  double *d_vals;
  B_TRY(T.get(&d_rows, nnz));
  B_TRY(T.get(&d_cols, nnz));
  B_TRY(T.get(&d_vals, nnz));
  CU_TRY(cudaMemcpyAsync(d_rows, h_rows, nnz * 4, cudaMemcpyHostToDevice, s));
  CU_TRY(cudaMemcpyAsync(d_cols, h_cols, nnz * 4, cudaMemcpyHostToDevice, s));
  CU_TRY(cudaMemcpyAsync(d_vals, h_vals, nnz * 8, cudaMemcpyHostToDevice, s));
  return coo_dev_to_plain(c, nnz, d_rows, d_cols, d_vals, sub, A, was_sorted);
}

// ---- text -> records on the device ------------------------------------------------
// One thread per line with the exact, strict parser of parse.cuh; lines it does
// not accept are listed for the host, which runs strtoul / strtod on them.
struct FlagToU64 {
  __host__ __device__ uint64_t operator()(uint8_t f) const { return f; }
};

__global__ void k_flag_newlines(const char *__restrict__ text, uint64_t len, uint8_t *flag) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < len)
    flag[i] = text[i] == '\n';
}

__global__ void k_parse_lines(uint64_t nnz, const char *__restrict__ text,
                              const uint64_t *__restrict__ nlpos, uint32_t *rows,
                              uint32_t *cols, double *vals, uint8_t *ask) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nnz)
    return;
  const char *p = text + (i ? nlpos[i - 1] + 1 : 0), *nl = text + nlpos[i];
  uint32_t r = 0, c = 0;
  double v = 0.0;
  ask[i] = (uint8_t)b2_parse_record(p, nl, r, c, v);
  rows[i] = r, cols[i] = c, vals[i] = v;
}

__global__ void k_patch_records(uint64_t n, const uint64_t *__restrict__ at,
                                const uint32_t *__restrict__ pr, const uint32_t *__restrict__ pc,
                                const double *__restrict__ pv, uint32_t *rows, uint32_t *cols,
                                double *vals) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    rows[at[i]] = pr[i], cols[at[i]] = pc[i], vals[at[i]] = pv[i];
}

// body: the text after the header line, `len` bytes, in host memory.  The first
// nnz lines are the records.  Returns B200_EINVAL (nothing else touched) when
// the text is not one strict record per line -- the caller then tokenises it
// the slow way, with the semantics of fscanf.
static int text_to_plain(b200_ctx *c, const char *body, uint64_t len, uint64_t nnz,
                         uint32_t sub, PlainCsr *A, uint64_t *n_host) {
  if (nnz == 0 || nnz >= 0xffffffffull)
    B_FAIL(B200_ERANGE, "ingest: nnz=%llu outside (0, 2^32)", (unsigned long long)nnz);
  if (len == 0)
    B_FAIL(B200_EINVAL, "ingest: empty matrix body");
  cudaStream_t s = c->stream;
  IngestTmp T;
  char *d_text;
  uint8_t *d_flag, *d_ask;
  uint64_t *d_pos, *d_count;
  B_TRY(T.get(&d_text, len));
  B_TRY(T.get(&d_flag, len));
  B_TRY(T.get(&d_count, 1));
  CU_TRY(cudaMemcpyAsync(d_text, body, len, cudaMemcpyHostToDevice, s));
  k_flag_newlines<<<nblk(len), T256, 0, s>>>(d_text, len, d_flag);
  {  // how many lines, then where they end
    void *tmp = nullptr;
    size_t bytes = 0;
    cub::TransformInputIterator<uint64_t, FlagToU64, const uint8_t *> wide(d_flag, FlagToU64());
    CU_TRY(cub::DeviceReduce::Sum(nullptr, bytes, wide, d_count, len, s));
    CU_TRY(cudaMalloc(&tmp, bytes ? bytes : 8));
    cudaError_t e = cub::DeviceReduce::Sum(tmp, bytes, wide, d_count, len, s);
    cudaStreamSynchronize(s);
    cudaFree(tmp);
    CU_TRY(e);
  }
  uint64_t nlines = 0;
  CU_TRY(cudaMemcpy(&nlines, d_count, 8, cudaMemcpyDeviceToHost));
  if (nlines < nnz)
    B_FAIL(B200_EINVAL, "ingest: %llu records announced, %llu lines", (unsigned long long)nnz,
           (unsigned long long)nlines);
  B_TRY(T.get(&d_pos, nlines + 1));
  {
    void *tmp = nullptr;
    size_t bytes = 0;
    cub::CountingInputIterator<uint64_t> it(0);
    CU_TRY(cub::DeviceSelect::Flagged(nullptr, bytes, it, d_flag, d_pos, d_count, len, s));
    CU_TRY(cudaMalloc(&tmp, bytes ? bytes : 8));
    cudaError_t e = cub::DeviceSelect::Flagged(tmp, bytes, it, d_flag, d_pos, d_count, len, s);
    cudaStreamSynchronize(s);
    cudaFree(tmp);
    CU_TRY(e);
  }
  uint32_t *d_rows, *d_cols;
  double *d_vals;
  B_TRY(T.get(&d_rows, nnz));
  B_TRY(T.get(&d_cols, nnz));
  B_TRY(T.get(&d_vals, nnz));
  B_TRY(T.get(&d_ask, nnz));
  k_parse_lines<<<nblk(nnz), T256, 0, s>>>(nnz, d_text, d_pos, d_rows, d_cols, d_vals, d_ask);
  CU_TRY(cudaGetLastError());
  // lines for the host
  uint64_t *d_list;
  B_TRY(T.get(&d_list, nnz));
  {
    void *tmp = nullptr;
    size_t bytes = 0;
    cub::CountingInputIterator<uint64_t> it(0);
    CU_TRY(cub::DeviceSelect::Flagged(nullptr, bytes, it, d_ask, d_list, d_count, nnz, s));
    CU_TRY(cudaMalloc(&tmp, bytes ? bytes : 8));
    cudaError_t e = cub::DeviceSelect::Flagged(tmp, bytes, it, d_ask, d_list, d_count, nnz, s);
    cudaStreamSynchronize(s);
    cudaFree(tmp);
    CU_TRY(e);
  }
  uint64_t nask = 0;
  CU_TRY(cudaMemcpy(&nask, d_count, 8, cudaMemcpyDeviceToHost));
  if (n_host)
    *n_host = nask;
  if (nask) {
    std::vector<uint64_t> list(nask), hpos(nnz);
    CU_TRY(cudaMemcpy(list.data(), d_list, nask * 8, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(hpos.data(), d_pos, nnz * 8, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> pr(nask), pc(nask);
    std::vector<double> pv(nask);
    for (uint64_t k = 0; k < nask; k++) {
      // [b, e]: the line, e at its newline
      const uint64_t b = list[k] ? hpos[list[k] - 1] + 1 : 0, e = hpos[list[k]];
      // libc on this one line; the value must end exactly at the newline
      const std::string line(body + b, body + e + 1);
      const char *p0 = line.c_str(), *last = p0 + line.size() - 1;
      char *q1, *q2, *q3;
      const unsigned long r = strtoul(p0, &q1, 10);
      const unsigned long cc = strtoul(q1, &q2, 10);
      const double v = strtod(q2, &q3);
      const bool strict = !(*p0 == ' ' || *p0 == '\t' || *p0 == '\n' || *p0 == '\r');
      if (!strict || q1 == p0 || q2 == q1 || q3 == q2 || q3 != last)
        B_FAIL(B200_EINVAL, "ingest: record %llu is not one strict line",
               (unsigned long long)list[k]);
      pr[k] = (uint32_t)r, pc[k] = (uint32_t)cc, pv[k] = v;
    }
    uint32_t *d_pr, *d_pc;
    double *d_pv;
    B_TRY(T.get(&d_pr, nask));
    B_TRY(T.get(&d_pc, nask));
    B_TRY(T.get(&d_pv, nask));
    CU_TRY(cudaMemcpyAsync(d_pr, pr.data(), nask * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d_pc, pc.data(), nask * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(d_pv, pv.data(), nask * 8, cudaMemcpyHostToDevice, s));
    k_patch_records<<<nblk(nask), T256, 0, s>>>(nask, d_list, d_pr, d_pc, d_pv, d_rows, d_cols, d_vals);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(s));
  }
  return coo_dev_to_plain(c, nnz, d_rows, d_cols, d_vals, sub, A, nullptr);
}

__global__ void k_narrow_offs(const uint64_t *in, uint32_t *out, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = (uint32_t)in[i];
}

extern "C" int b200_coo_to_csr(b200_ctx *c, uint64_t nnz, const uint32_t *rows,
                               const uint32_t *cols, const double *vals,
                               uint32_t *nrows_out, uint64_t *nnz_out,
                               uint32_t *offs, uint32_t *cols_out,
                               double *vals_out) {
  if (!c || !rows || !cols || !vals || !nrows_out || !nnz_out)
    B_FAIL(B200_EINVAL, "b200_coo_to_csr: null argument");
  CU_TRY(cudaSetDevice(c->device));
  PlainCsr A;
  int rc = coo_to_plain(c, nnz, rows, cols, vals, 0, &A, nullptr);
  if (rc != B200_OK) {
    plain_free(&A);
    return rc;
  }
  *nrows_out = (uint32_t)A.n, *nnz_out = A.nnz;
  if (offs && cols_out && vals_out) {
    cudaStream_t s = c->stream;
    uint32_t *o32 = nullptr;
    CU_TRY(cudaMalloc(&o32, (A.n + 1) * 4));
    k_narrow_offs<<<nblk(A.n + 1), T256, 0, s>>>(A.offs, o32, A.n + 1);
    CU_TRY(cudaMemcpyAsync(offs, o32, (A.n + 1) * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(cols_out, A.cols, A.nnz * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(vals_out, A.vals, A.nnz * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    cudaFree(o32);
  }
  plain_free(&A);
  return B200_OK;
}

static int plain_to_host(b200_ctx *c, PlainCsr &A, uint32_t *nrows_out, uint64_t *nnz_out,
                         uint32_t *offs, uint32_t *cols_out, double *vals_out) {
  *nrows_out = (uint32_t)A.n, *nnz_out = A.nnz;
  if (offs && cols_out && vals_out) {
    cudaStream_t s = c->stream;
    uint32_t *o32 = nullptr;
    CU_TRY(cudaMalloc(&o32, (A.n + 1) * 4));
    k_narrow_offs<<<nblk(A.n + 1), T256, 0, s>>>(A.offs, o32, A.n + 1);
    CU_TRY(cudaMemcpyAsync(offs, o32, (A.n + 1) * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(cols_out, A.cols, A.nnz * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(vals_out, A.vals, A.nnz * 8, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    cudaFree(o32);
  }
  return B200_OK;
}

extern "C" int b200_text_to_csr(b200_ctx *c, const char *body, uint64_t len, uint64_t nnz,
                                uint32_t *nrows_out, uint64_t *nnz_out, uint32_t *offs,
                                uint32_t *cols_out, double *vals_out, uint64_t *n_host_parsed) {
  if (!c || !body || !nrows_out || !nnz_out)
    B_FAIL(B200_EINVAL, "b200_text_to_csr: null argument");
  CU_TRY(cudaSetDevice(c->device));
  PlainCsr A;
  int rc = text_to_plain(c, body, len, nnz, 0, &A, n_host_parsed);
  if (rc == B200_OK)
    rc = plain_to_host(c, A, nrows_out, nnz_out, offs, cols_out, vals_out);
  plain_free(&A);
  return rc;
}

int mat_from_plain(b200_ctx *c, PlainCsr *A, uint32_t flags, b200_mat **out);

extern "C" int b200_mat_from_coo(b200_ctx *c, uint64_t nnz, uint32_t base,
                                 const uint32_t *rows, const uint32_t *cols,
                                 const double *vals, uint32_t flags,
                                 b200_mat **out) {
  if (!c || !rows || !cols || !vals || !out)
    B_FAIL(B200_EINVAL, "b200_mat_from_coo: null argument");
  if (base > 1)
    B_FAIL(B200_EINVAL, "b200_mat_from_coo: base=%u", base);
  if ((flags & B200_MAT_FORCE_SELL) && (flags & B200_MAT_FORCE_VECTOR))
    B_FAIL(B200_EINVAL, "b200_mat_from_coo: contradictory FORCE flags");
  CU_TRY(cudaSetDevice(c->device));
  *out = nullptr;
  PlainCsr A;
  int rc = coo_to_plain(c, nnz, rows, cols, vals, base, &A, nullptr);
  if (rc != B200_OK) {
    plain_free(&A);
    return rc;
  }
  return mat_from_plain(c, &A, flags, out);
}
