// dist.cu -- north_star piece (5): row-block partition over the ranks, halo
// exchange overlapped with the interior SpMV, and the all-reduce of the CG
// scalars.  One rank = one process (or host thread) = one GPU.
//
// No reference counterpart: every lsbench backend is single-device
// (src/amgx.c:88-89, src/ginkgo.cpp:18, src/hypre.c:31).
//
//   partition   contiguous row blocks [cut(k), cut(k+1)), cuts rounded to 32
//               rows (whole z-planes for the stencils when P divides N).
//   renumber    owned columns -> [0, n_local); remote columns -> n_local + slot,
//               slots in ascending global order (so grouped by owner).  The
//               entry order inside a row is untouched: a row sums in the same
//               order on 1 GPU and on P.
//   interior    maximal middle run of rows without remote columns; it is
//               multiplied while the halo is in flight on the comm stream.
//   exchange    pack kernel + grouped ncclSend/ncclRecv over NVLink.
//   scalars     1-3 fp64 sums per reduction: written by the producing kernel's last
//               CTA into every rank's mailbox over NVLink peer memory and summed
//               by the consuming kernel (xr_setup here, xr_push / xr_wait_sum in
//               common.cuh); ncclAllReduce when peer memory cannot be mapped.
//
// NCCL is bound with dlopen so the library loads (and the single-GPU path
// runs) on hosts without it, and so that inside a torch process the NCCL that
// torch already loaded is the one used.
#include "common.cuh"
#include <cub/cub.cuh>
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>
#include <vector>

#define T256 256
static inline unsigned nblk(uint64_t n, unsigned t = T256) {
  return (unsigned)((n + t - 1) / t);
}


struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t,
                            ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t,
                            ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t,
                       cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t,
                       cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char *(*GetErrorString)(ncclResult_t);
};

static NcclApi g_nccl;
static bool g_nccl_ok = false;

static int nccl_bind() {
  if (g_nccl_ok)
    return B200_OK;
  void *h = nullptr;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names)
    if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL)))
      break;
  if (!h)
    B_FAIL(B200_ENCCL, "b200: cannot dlopen libnccl.so.2: %s", dlerror());
#define BIND(field, sym)                                                       \
  if (!(*(void **)(&g_nccl.field) = dlsym(h, sym)))                            \
    B_FAIL(B200_ENCCL, "b200: libnccl lacks %s", sym);
  BIND(GetUniqueId, "ncclGetUniqueId")
  BIND(CommInitRank, "ncclCommInitRank")
  BIND(CommDestroy, "ncclCommDestroy")
  BIND(AllReduce, "ncclAllReduce")
  BIND(AllGather, "ncclAllGather")
  BIND(Send, "ncclSend")
  BIND(Recv, "ncclRecv")
  BIND(GroupStart, "ncclGroupStart")
  BIND(GroupEnd, "ncclGroupEnd")
  BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
  g_nccl_ok = true;
  return B200_OK;
}

#define NC_TRY(expr)                                                           \
  do {                                                                         \
    ncclResult_t r_ = (expr);                                                  \
    if (r_ != ncclSuccess) {                                                   \
      b200_set_error("%s:%d nccl error: %s (%s)", __FILE__, __LINE__,          \
                     g_nccl.GetErrorString(r_), #expr);                        \
      return B200_ENCCL;                                                       \
    }                                                                          \
  } while (0)

extern "C" int b200_nccl_unique_id(void *id_out) {
  if (!id_out)
    B_FAIL(B200_EINVAL, "b200_nccl_unique_id: null argument");
  B_TRY(nccl_bind());
  static_assert(sizeof(ncclUniqueId) == B200_NCCL_ID_BYTES, "nccl id size");
  NC_TRY(g_nccl.GetUniqueId((ncclUniqueId *)id_out));
  return B200_OK;
}

// ---- peer-memory all-reduce: mailboxes every rank can write (common.cuh) -----------
// Called once per context after the communicator exists.  All ranks must end up
// on the same path, so the decision is itself all-reduced (min).  Not an error
// when peer memory cannot be mapped: the NCCL all-reduce stays in use.
// B200_ALLREDUCE=nccl switches it off.
struct XrInfo {
  long long pid;
  int dev, ok;
  unsigned long long ptr;
  cudaIpcMemHandle_t handle;
};

static int xr_setup(b200_ctx *c) {
  const char *e = getenv("B200_ALLREDUCE");
  int want = !(e && strcmp(e, "nccl") == 0) && c->nranks <= B2_XR_MAX_RANKS;
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  cudaStream_t s = c->stream;
  const size_t mail_bytes = (size_t)B2_XR_KINDS * B2_XR_MAX_RANKS * 4 * sizeof(double);
  XrInfo mine;
  memset(&mine, 0, sizeof mine);
  mine.pid = (long long)getpid(), mine.dev = c->device, mine.ok = want;
  if (want) {
    CU_TRY(cudaMalloc(&c->xr_mail, mail_bytes));
    CU_TRY(cudaMemsetAsync(c->xr_mail, 0, mail_bytes, s));
    mine.ptr = (unsigned long long)c->xr_mail;
    if (cudaIpcGetMemHandle(&mine.handle, c->xr_mail) != cudaSuccess)
      cudaGetLastError(), mine.ok = 0;
  }
  // everybody's record to everybody
  XrInfo *d_all = nullptr;
  std::vector<XrInfo> all(c->nranks);
  CU_TRY(cudaMalloc(&d_all, sizeof(XrInfo) * c->nranks));
  CU_TRY(cudaMemcpyAsync(d_all + c->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, s));
  NC_TRY(g_nccl.AllGather(d_all + c->rank, d_all, sizeof(XrInfo), ncclChar, comm, s));
  CU_TRY(cudaMemcpyAsync(all.data(), d_all, sizeof(XrInfo) * c->nranks, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  std::vector<double *> peers(B2_XR_MAX_RANKS, nullptr);
  int ok = 1;
  for (int r = 0; r < c->nranks && ok; r++) {
    if (!all[r].ok) {
      ok = 0;
    } else if (r == c->rank) {
      peers[r] = c->xr_mail;
    } else if (all[r].pid == mine.pid) {  // another host thread of this process
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, c->device, all[r].dev) != cudaSuccess || !can) {
        cudaGetLastError(), ok = 0;
        break;
      }
      cudaError_t pe = cudaDeviceEnablePeerAccess(all[r].dev, 0);
      if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
        ok = 0;
      cudaGetLastError();
      peers[r] = (double *)all[r].ptr;
    } else {
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError(), ok = 0;
        break;
      }
      c->xr_opened[r] = p, peers[r] = (double *)p;
    }
  }
  // one decision for all ranks
  int *d_ok = (int *)d_all;
  CU_TRY(cudaMemcpyAsync(d_ok, &ok, sizeof ok, cudaMemcpyHostToDevice, s));
  NC_TRY(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm, s));
  CU_TRY(cudaMemcpyAsync(&ok, d_ok, sizeof ok, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  cudaFree(d_all);
  if (ok) {
    for (int r = 0; r < B2_XR_MAX_RANKS; r++)
      c->xr_peers_h[r] = peers[r];
    CU_TRY(cudaMalloc(&c->d_seq, 2 * sizeof(unsigned long long)));
    CU_TRY(cudaMemsetAsync(c->d_seq, 0, 2 * sizeof(unsigned long long), s));
    CU_TRY(cudaMalloc(&c->xr_peers, sizeof(double *) * B2_XR_MAX_RANKS));
    CU_TRY(cudaMemcpyAsync(c->xr_peers, peers.data(), sizeof(double *) * B2_XR_MAX_RANKS,
                           cudaMemcpyHostToDevice, s));
    CU_TRY(cudaStreamSynchronize(s));
  }
  c->xr_on = ok != 0;
  if (getenv("B200_VERBOSE") && c->rank == 0)
    fprintf(stderr, "b200: CG scalars all-reduced over %s\n",
            c->xr_on ? "peer memory (fused into producer / consumer kernels)" : "NCCL");
  return B200_OK;
}

int dist_comm_init(b200_ctx *c, const void *nccl_id) {
  B_TRY(nccl_bind());
  ncclUniqueId id;
  memcpy(&id, nccl_id, sizeof id);
  ncclComm_t comm;
  NC_TRY(g_nccl.CommInitRank(&comm, c->nranks, id, c->rank));
  c->nccl_comm = comm, c->nccl = &g_nccl;
  return xr_setup(c);
}

void dist_comm_destroy(b200_ctx *c) {
  for (void *&p : c->xr_opened)
    if (p)
      cudaIpcCloseMemHandle(p), p = nullptr;
  if (c->xr_peers) cudaFree(c->xr_peers);
  if (c->xr_mail) cudaFree(c->xr_mail);
  if (c->d_seq) cudaFree(c->d_seq);
  c->d_seq = nullptr;
  c->xr_peers = nullptr, c->xr_mail = nullptr, c->xr_on = false;
  if (c->nccl_comm)
    g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
  c->nccl_comm = nullptr;
}

int allreduce_sum(b200_ctx *c, const double *src, double *dst, int n) {
  if (c->nranks == 1)
    return B200_OK;
  NC_TRY(g_nccl.AllReduce(src, dst, n, ncclDouble, ncclSum,
                          (ncclComm_t)c->nccl_comm, c->stream));
  return B200_OK;
}

// ---- partition / renumber ---------------------------------------------------------
__global__ void k_mark_remote(uint64_t nloc, const uint64_t *offs,
                              const uint32_t *cols, uint64_t r0, uint64_t r1,
                              uint32_t *bits, unsigned long long *lo_max,
                              unsigned long long *hi_min) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nloc)
    return;
  bool remote = false;
  for (uint64_t k = offs[i]; k < offs[i + 1]; k++) {
    uint64_t c = cols[k];
    if (c < r0 || c >= r1) {
      atomicOr(&bits[c >> 5], 1u << (c & 31));
      remote = true;
    }
  }
  if (remote) {
    // interior = the middle run between the last boundary row of the first
    // half and the first boundary row of the second half
    if (i < nloc / 2)
      atomicMax(lo_max, (unsigned long long)(i + 1));
    else
      atomicMin(hi_min, (unsigned long long)i);
  }
}

__global__ void k_popc(const uint32_t *bits, uint64_t nwords, uint32_t *cnt) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < nwords)
    cnt[i] = __popc(bits[i]);
  if (i == nwords)
    cnt[i] = 0;
}

__global__ void k_emit_halo(const uint32_t *bits, uint64_t nwords,
                            const uint32_t *rank_of_word, uint64_t *gcols) {
  uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (w >= nwords)
    return;
  uint32_t b = bits[w];
  uint64_t o = rank_of_word[w];
  while (b) {
    int bit = __ffs(b) - 1;
    gcols[o++] = w * 32 + bit;
    b &= b - 1;
  }
}

__global__ void k_renumber(uint64_t nnz, uint32_t *cols, uint64_t r0, uint64_t r1,
                           uint64_t nloc, const uint32_t *bits,
                           const uint32_t *rank_of_word) {
  uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (k >= nnz)
    return;
  uint64_t c = cols[k];
  if (c >= r0 && c < r1)
    cols[k] = (uint32_t)(c - r0);
  else
    cols[k] = (uint32_t)(nloc + rank_of_word[c >> 5] +
                         __popc(bits[c >> 5] & ((1u << (c & 31)) - 1u)));
}

__global__ void k_shift_offs(const uint64_t *in, uint64_t *out, uint64_t n,
                             uint64_t base) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    out[i] = in[i] - base;
}

// A holds rows [row_begin, row_begin + A->n) -- or, when it holds the whole
// matrix on a multi-rank context, it is cut down to this rank's block first.
int partition_and_renumber(b200_ctx *c, PlainCsr *A, uint64_t n_global,
                           uint64_t row_begin, b200_mat *M) {
  M->row_begin = row_begin;
  if (c->nranks == 1)
    return B200_OK;
  cudaStream_t s = c->stream;
  uint64_t r0, r1;
  b200_row_block(n_global, c->rank, c->nranks, &r0, &r1);
  if (A->n == n_global && !(r0 == 0 && r1 == n_global)) {
    // cut the block out of the full matrix
    uint64_t nloc = r1 - r0, k0 = 0, k1 = 0;
    CU_TRY(cudaMemcpy(&k0, A->offs + r0, 8, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(&k1, A->offs + r1, 8, cudaMemcpyDeviceToHost));
    PlainCsr B;
    B.n = nloc, B.nnz = k1 - k0;
    CU_TRY(cudaMalloc(&B.offs, (nloc + 1) * 8));
    CU_TRY(cudaMalloc(&B.cols, (B.nnz ? B.nnz : 1) * 4));
    CU_TRY(cudaMalloc(&B.vals, (B.nnz ? B.nnz : 1) * 8));
    k_shift_offs<<<nblk(nloc + 1), T256, 0, s>>>(A->offs + r0, B.offs, nloc + 1, k0);
    CU_TRY(cudaMemcpyAsync(B.cols, A->cols + k0, B.nnz * 4, cudaMemcpyDeviceToDevice, s));
    CU_TRY(cudaMemcpyAsync(B.vals, A->vals + k0, B.nnz * 8, cudaMemcpyDeviceToDevice, s));
    CU_TRY(cudaStreamSynchronize(s));
    plain_free(A);
    *A = B;
  } else if (row_begin != r0 || A->n != r1 - r0) {
    B_FAIL(B200_EINVAL, "partition: rows [%llu,+%llu) are not this rank's block",
           (unsigned long long)row_begin, (unsigned long long)A->n);
  }
  M->row_begin = r0;
  const uint64_t nloc = A->n, nwords = (n_global + 31) / 32;
  uint32_t *bits, *cnt, *rank_of_word;
  unsigned long long *d_lohi, h_lohi[2] = {0ull, (unsigned long long)nloc};
  CU_TRY(cudaMalloc(&bits, (nwords + 1) * 4));
  CU_TRY(cudaMalloc(&cnt, (nwords + 1) * 4));
  CU_TRY(cudaMalloc(&rank_of_word, (nwords + 1) * 4));
  CU_TRY(cudaMalloc(&d_lohi, 16));
  CU_TRY(cudaMemsetAsync(bits, 0, (nwords + 1) * 4, s));
  CU_TRY(cudaMemcpyAsync(d_lohi, h_lohi, 16, cudaMemcpyHostToDevice, s));
  if (nloc)
    k_mark_remote<<<nblk(nloc), T256, 0, s>>>(nloc, A->offs, A->cols, r0, r1, bits,
                                              d_lohi, d_lohi + 1);
  k_popc<<<nblk(nwords + 1), T256, 0, s>>>(bits, nwords, cnt);
  {
    void *tmp = nullptr;
    size_t bytes = 0;
    CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, rank_of_word, nwords + 1, s));
    CU_TRY(cudaMalloc(&tmp, bytes ? bytes : 8));
    CU_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, cnt, rank_of_word, nwords + 1, s));
    CU_TRY(cudaStreamSynchronize(s));
    cudaFree(tmp);
  }
  uint32_t n_halo = 0;
  CU_TRY(cudaMemcpy(&n_halo, rank_of_word + nwords, 4, cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(h_lohi, d_lohi, 16, cudaMemcpyDeviceToHost));
  if (nloc + (uint64_t)n_halo >= 0xffffffffull)
    B_FAIL(B200_ERANGE, "partition: local + halo columns exceed 32 bits");
  M->halo.n_halo = n_halo;
  {  // (row-block cuts are multiples of 32: the word of r0 starts at r0)
    uint32_t below = 0;
    CU_TRY(cudaMemcpy(&below, rank_of_word + r0 / 32, 4, cudaMemcpyDeviceToHost));
    M->halo.n_low = below;
  }
  CU_TRY(cudaMalloc(&M->halo.d_gcols, (n_halo + 1ull) * 8));
  k_emit_halo<<<nblk(nwords), T256, 0, s>>>(bits, nwords, rank_of_word, M->halo.d_gcols);
  if (A->nnz)
    k_renumber<<<nblk(A->nnz), T256, 0, s>>>(A->nnz, A->cols, r0, r1, nloc, bits,
                                             rank_of_word);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(s));
  M->interior_begin = h_lohi[0];
  M->interior_end = h_lohi[1] > h_lohi[0] ? h_lohi[1] : h_lohi[0];
  cudaFree(bits), cudaFree(cnt), cudaFree(rank_of_word), cudaFree(d_lohi);
  return B200_OK;
}

// ---- halo plan ----------------------------------------------------------------------
__global__ void k_global_to_local_u32(const uint64_t *g, uint32_t *l, uint64_t n,
                                      uint64_t r0) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    l[i] = (uint32_t)(g[i] - r0);
}

int halo_setup(b200_mat *M) {
  b200_ctx *c = M->ctx;
  HaloPlan &H = M->halo;
  const int P = c->nranks, me = c->rank;
  cudaStream_t s = c->stream;
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  // how many of my halo slots each owner holds (slots are sorted by global id)
  uint64_t *h_g = (uint64_t *)malloc((H.n_halo + 1) * 8);
  CU_TRY(cudaMemcpy(h_g, H.d_gcols, H.n_halo * 8, cudaMemcpyDeviceToHost));
  uint64_t *need = (uint64_t *)calloc((size_t)P * P, 8);  // need[r*P + o]
  uint64_t *first = (uint64_t *)calloc(P + 1, 8);
  {
    uint64_t j = 0;
    for (int o = 0; o < P; o++) {
      uint64_t b0, b1;
      b200_row_block(M->n_global, o, P, &b0, &b1);
      first[o] = j;
      while (j < H.n_halo && h_g[j] < b1)
        j++;
      need[(size_t)me * P + o] = j - first[o];
    }
    first[P] = j;
  }
  free(h_g);
  uint64_t *d_need;
  CU_TRY(cudaMalloc(&d_need, (size_t)P * P * 8));
  CU_TRY(cudaMemcpyAsync(d_need + (size_t)me * P, need + (size_t)me * P, P * 8,
                         cudaMemcpyHostToDevice, s));
  NC_TRY(g_nccl.AllGather(d_need + (size_t)me * P, d_need, P, ncclUint64, comm, s));
  CU_TRY(cudaStreamSynchronize(s));
  CU_TRY(cudaMemcpy(need, d_need, (size_t)P * P * 8, cudaMemcpyDeviceToHost));
  cudaFree(d_need);

  H.peer = (int *)calloc(P, sizeof(int));
  H.recv_off = (uint64_t *)calloc(P + 1, 8);
  H.send_off = (uint64_t *)calloc(P + 1, 8);
  H.n_peers = 0;
  uint64_t ns = 0;
  for (int o = 0; o < P; o++) {
    uint64_t rcv = need[(size_t)me * P + o], snd = need[(size_t)o * P + me];
    if (o == me || (rcv == 0 && snd == 0))
      continue;
    int k = H.n_peers++;
    H.peer[k] = o;
    H.recv_off[k] = first[o];
    H.send_off[k] = ns;
    ns += snd;
    // lengths are recovered from need[] below; store ends in the next slot
    H.recv_off[k + 1] = first[o] + rcv;
    H.send_off[k + 1] = ns;
  }
  H.n_send = ns;
  // recv ranges are not necessarily adjacent across skipped owners: keep
  // explicit (begin, count) pairs
  uint64_t *rb = (uint64_t *)calloc(2 * (size_t)P + 2, 8);
  uint64_t *sb = (uint64_t *)calloc(2 * (size_t)P + 2, 8);
  {
    uint64_t acc = 0;
    for (int k = 0; k < H.n_peers; k++) {
      int o = H.peer[k];
      rb[2 * k] = first[o], rb[2 * k + 1] = need[(size_t)me * P + o];
      sb[2 * k] = acc, sb[2 * k + 1] = need[(size_t)o * P + me];
      acc += sb[2 * k + 1];
    }
  }
  free(H.recv_off), free(H.send_off);
  H.recv_off = rb, H.send_off = sb;

  // tell every owner which of its rows I read
  uint64_t *d_send_g;
  CU_TRY(cudaMalloc(&d_send_g, (ns + 1) * 8));
  NC_TRY(g_nccl.GroupStart());
  for (int k = 0; k < H.n_peers; k++) {
    if (rb[2 * k + 1])
      NC_TRY(g_nccl.Send(H.d_gcols + rb[2 * k], rb[2 * k + 1], ncclUint64, H.peer[k], comm, s));
    if (sb[2 * k + 1])
      NC_TRY(g_nccl.Recv(d_send_g + sb[2 * k], sb[2 * k + 1], ncclUint64, H.peer[k], comm, s));
  }
  NC_TRY(g_nccl.GroupEnd());
  CU_TRY(cudaMalloc(&H.d_send_idx, (ns + 1) * 4));
  CU_TRY(cudaMalloc(&H.d_send_buf, (ns + 1) * 8));
  if (ns)
    k_global_to_local_u32<<<nblk(ns), T256, 0, s>>>(d_send_g, H.d_send_idx, ns, M->row_begin);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(s));
  cudaFree(d_send_g);
  M->device_bytes += (ns + 1) * 12 + (H.n_halo + 1) * 8;
  free(need), free(first);
  return B200_OK;
}

void halo_free(b200_mat *M) {
  HaloPlan &H = M->halo;
  for (void *&p : H.peer_opened)
    if (p)
      cudaIpcCloseMemHandle(p), p = nullptr;
  if (H.d_push_ticket) cudaFree(H.d_push_ticket);
  if (H.d_gcols) cudaFree(H.d_gcols);
  if (H.d_send_idx) cudaFree(H.d_send_idx);
  if (H.d_send_buf) cudaFree(H.d_send_buf);
  free(H.peer), free(H.recv_off), free(H.send_off);
  H = HaloPlan();
}

__global__ void k_halo_pack(const double *__restrict__ x,
                            const uint32_t *__restrict__ idx,
                            double *__restrict__ buf, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n)
    buf[i] = x[idx[i]];
}

// Owned part of x_ext is final on the compute stream; ship the entries the
// neighbours read and receive mine into x_ext[n_local ...], all on the comm
// stream so the interior SpMV runs meanwhile.
int halo_exchange_begin(b200_mat *M, double *x_ext) {
  b200_ctx *c = M->ctx;
  if (c->nranks == 1)
    return B200_OK;
  HaloPlan &H = M->halo;
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  CU_TRY(cudaEventRecord(c->ev_ready, c->stream));
  CU_TRY(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
  if (H.n_send)
    k_halo_pack<<<nblk(H.n_send), T256, 0, c->comm_stream>>>(x_ext, H.d_send_idx,
                                                             H.d_send_buf, H.n_send);
  c->launches += H.n_send > 0;
  CU_TRY(cudaGetLastError());
  NC_TRY(g_nccl.GroupStart());
  for (int k = 0; k < H.n_peers; k++) {
    if (H.send_off[2 * k + 1])
      NC_TRY(g_nccl.Send(H.d_send_buf + H.send_off[2 * k], H.send_off[2 * k + 1],
                         ncclDouble, H.peer[k], comm, c->comm_stream));
    if (H.recv_off[2 * k + 1])
      NC_TRY(g_nccl.Recv(x_ext + M->n_local + H.recv_off[2 * k], H.recv_off[2 * k + 1],
                         ncclDouble, H.peer[k], comm, c->comm_stream));
  }
  NC_TRY(g_nccl.GroupEnd());
  CU_TRY(cudaEventRecord(c->ev_halo, c->comm_stream));
  return B200_OK;
}

int halo_exchange_wait(b200_mat *M) {
  b200_ctx *c = M->ctx;
  if (c->nranks == 1)
    return B200_OK;
  CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
  return B200_OK;
}

// ---- halo exchange over peer memory -----------------------------------------------------
// Inside the PCG iteration the halo of p does not go through NCCL: a small kernel
// stores the rows the neighbours read STRAIGHT into the halo section of their p
// vector over NVLink (the vector is mapped here: peer access between host threads,
// CUDA IPC between processes), then -- last CTA, after a system-scope fence -- this
// iteration's sequence number into their mailbox (kind 2).  The neighbour's boundary
// rows start behind a one-warp kernel that waits for the sequence number of everyone
// it reads from.  No collective call, nothing the host has to do per iteration: the
// chunk of iterations is a CUDA graph on several ranks as well.
//
// Why no neighbour can overwrite a halo that is still being read: rank A pushes p of
// iteration k+1 after its K3(k), which needs {r.z, r.r}(k) from every rank, which rank
// B sends from its K2(k) -- behind B's boundary rows of iteration k in B's stream.
// That chain only exists inside the iteration; the stand-alone SpMV, the start-up
// and the exit products keep the NCCL exchange.
struct PeerHaloInfo {
  long long pid;
  int dev, ok;
  unsigned long long ptr;      // w_p
  unsigned long long n_local;
  unsigned long long first[B2_XR_MAX_RANKS];  // where owner o's entries begin in my halo section
  cudaIpcMemHandle_t handle;
};

int halo_peer_setup(b200_mat *M) {
  b200_ctx *c = M->ctx;
  HaloPlan &H = M->halo;
  if (H.peer_tried || c->nranks == 1)
    return B200_OK;
  H.peer_tried = true;
  const char *e = getenv("B200_HALO");
  const int want = c->xr_on && !(e && strcmp(e, "nccl") == 0);
  ncclComm_t comm = (ncclComm_t)c->nccl_comm;
  cudaStream_t s = c->stream;
  const int P = c->nranks, me = c->rank;
  PeerHaloInfo mine;
  memset(&mine, 0, sizeof mine);
  mine.pid = (long long)getpid(), mine.dev = c->device, mine.ok = want && M->w_p != nullptr;
  mine.ptr = (unsigned long long)M->w_p, mine.n_local = M->n_local;
  for (int k = 0; k < H.n_peers; k++)
    mine.first[H.peer[k]] = H.recv_off[2 * k];
  if (mine.ok && cudaIpcGetMemHandle(&mine.handle, M->w_p) != cudaSuccess)
    cudaGetLastError(), mine.ok = 0;
  PeerHaloInfo *d_all = nullptr;
  std::vector<PeerHaloInfo> all(P);
  CU_TRY(cudaMalloc(&d_all, sizeof(PeerHaloInfo) * P));
  CU_TRY(cudaMemcpyAsync(d_all + me, &mine, sizeof mine, cudaMemcpyHostToDevice, s));
  NC_TRY(g_nccl.AllGather(d_all + me, d_all, sizeof(PeerHaloInfo), ncclChar, comm, s));
  CU_TRY(cudaMemcpyAsync(all.data(), d_all, sizeof(PeerHaloInfo) * P, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  int ok = 1;
  for (int r = 0; r < P; r++)
    ok &= all[r].ok;
  memset(&H.push, 0, sizeof H.push);
  memset(&H.wait, 0, sizeof H.wait);
  for (int k = 0; k < H.n_peers && ok; k++) {
    const int o = H.peer[k];
    if (H.recv_off[2 * k + 1])
      H.wait.src[H.wait.n_src++] = o;
    if (!H.send_off[2 * k + 1])
      continue;
    double *base = nullptr;
    if (all[o].pid == mine.pid) {
      base = (double *)all[o].ptr;  // peer access was enabled for the mailboxes (xr_setup)
    } else {
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[o].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError(), ok = 0;
        break;
      }
      H.peer_opened[o] = p, base = (double *)p;
    }
    const int j = H.push.n_dst++;
    H.push.seg_begin[j] = (uint32_t)H.send_off[2 * k];
    H.push.seg_begin[j + 1] = (uint32_t)(H.send_off[2 * k] + H.send_off[2 * k + 1]);
    H.push.dst[j] = base + all[o].n_local + all[o].first[me];
    H.push.flag[j] = reinterpret_cast<unsigned long long *>(
                         c->xr_peers_h[o] + ((size_t)2 * B2_XR_MAX_RANKS + me) * 4) + 3;
  }
  // (the send list is grouped by neighbour in the order of H.peer, so the runs are adjacent)
  int *d_ok = (int *)d_all;
  CU_TRY(cudaMemcpyAsync(d_ok, &ok, sizeof ok, cudaMemcpyHostToDevice, s));
  NC_TRY(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm, s));
  CU_TRY(cudaMemcpyAsync(&ok, d_ok, sizeof ok, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  cudaFree(d_all);
  if (ok) {
    CU_TRY(cudaMalloc(&H.d_push_ticket, sizeof(unsigned)));
    CU_TRY(cudaMemsetAsync(H.d_push_ticket, 0, sizeof(unsigned), s));
  }
  H.peer_ready = ok != 0;
  if (getenv("B200_VERBOSE") && me == 0)
    fprintf(stderr, "b200: halo of p exchanged over %s\n",
            H.peer_ready ? "peer memory (stores into the neighbours' vectors)" : "NCCL send/recv");
  return B200_OK;
}

__global__ void k_xr_chunk_begin(unsigned long long *seq, unsigned count) {
  seq[0] = seq[1];
  seq[1] += count;
}

int xr_chunk_begin(b200_ctx *c, unsigned count) {
  if (!c->d_seq)
    return B200_OK;
  k_xr_chunk_begin<<<1, 1, 0, c->stream>>>(c->d_seq, count);
  c->launches += 1;
  CU_TRY(cudaGetLastError());
  return B200_OK;
}

#define HP_THREADS 256
__global__ void __launch_bounds__(HP_THREADS)
k_halo_push(const double *__restrict__ x, const uint32_t *__restrict__ idx, uint32_t n_send,
            const HaloPushArgs a, const int *done, const unsigned long long *seq_base,
            unsigned seq_off, unsigned *ticket) {
  if (done && *done)
    return;
  __shared__ bool is_last;
  for (uint32_t i = blockIdx.x * HP_THREADS + threadIdx.x; i < n_send; i += gridDim.x * HP_THREADS) {
    int k = 0;
    while (k + 1 < a.n_dst && i >= a.seg_begin[k + 1])
      k++;
    a.dst[k][i - a.seg_begin[k]] = x[idx[i]];
  }
  __threadfence_system();  // my stores, before the ticket says this CTA is through
  __syncthreads();
  if (threadIdx.x == 0)
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last)
    return;
  if ((int)threadIdx.x < a.n_dst) {
    __threadfence_system();
    st_relaxed_sys(a.flag[threadIdx.x], *seq_base + seq_off);
  }
  if (threadIdx.x == 0)
    *ticket = 0;
}

__global__ void k_halo_wait(const double *mine, const HaloWaitArgs a, const unsigned long long *seq_base,
                            unsigned seq_off, PcgState *st) {
  if (st && st->done)
    return;
  const int k = threadIdx.x;
  if (k >= a.n_src)
    return;
  const unsigned long long want = *seq_base + seq_off;
  const unsigned long long *flag =
      reinterpret_cast<const unsigned long long *>(mine + ((size_t)2 * B2_XR_MAX_RANKS + a.src[k]) * 4) + 3;
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) != want)
    if (clock64() - t0 > 8000000000ll) {  // ~4 s: a neighbour is gone; stop, do not hang the GPU
      if (st)
        st->done = 1, st->status = 3;
      break;
    }
}

int halo_peer_push(b200_mat *M, const double *x_ext, unsigned seq_off) {
  b200_ctx *c = M->ctx;
  HaloPlan &H = M->halo;
  CU_TRY(cudaEventRecord(c->ev_ready, c->stream));
  CU_TRY(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
  if (H.n_send && H.push.n_dst) {
    unsigned g = nblk(H.n_send, HP_THREADS);
    g = g > 64u ? 64u : g;
    k_halo_push<<<g, HP_THREADS, 0, c->comm_stream>>>(x_ext, H.d_send_idx, (uint32_t)H.n_send, H.push,
                                                     M->state ? &M->state->done : nullptr, c->d_seq,
                                                     seq_off, H.d_push_ticket);
    c->launches += 1;
    CU_TRY(cudaGetLastError());
  }
  CU_TRY(cudaEventRecord(c->ev_halo, c->comm_stream));
  return B200_OK;
}

int halo_peer_wait(b200_mat *M, unsigned seq_off) {
  b200_ctx *c = M->ctx;
  HaloPlan &H = M->halo;
  CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_halo, 0));  // my own push read p: before p changes
  if (H.wait.n_src) {
    k_halo_wait<<<1, 32, 0, c->stream>>>(c->xr_mail, H.wait, c->d_seq, seq_off, M->state);
    c->launches += 1;
    CU_TRY(cudaGetLastError());
  }
  return B200_OK;
}

extern "C" int b200_mat_halo_cols(const b200_mat *M, uint64_t *g) {
  if (!M || !g)
    B_FAIL(B200_EINVAL, "b200_mat_halo_cols: null argument");
  if (M->halo.n_halo)
    CU_TRY(cudaMemcpy(g, M->halo.d_gcols, M->halo.n_halo * 8, cudaMemcpyDeviceToHost));
  return B200_OK;
}
