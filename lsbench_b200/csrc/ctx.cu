// ctx.cu -- context, error reporting and raw memory entry points of the C ABI.
// Replaces the file-scope handles a reference backend keeps
// (src/cusparse.c:33-36, created in cusparse_init :138-151).
#include "common.cuh"

static thread_local char g_err[512] = "";

void b200_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

extern "C" const char *b200_last_error(void) { return g_err; }
extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }

extern "C" int b200_device_count(int *count) {
  if (!count)
    B_FAIL(B200_EINVAL, "b200_device_count: null argument");
  CU_TRY(cudaGetDeviceCount(count));
  return B200_OK;
}

static int ctx_common(int device, b200_ctx **out) {
  if (!out)
    B_FAIL(B200_EINVAL, "b200_ctx_create: null out pointer");
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev)
    B_FAIL(B200_EINVAL, "b200_ctx_create: device %d of %d", device, ndev);
  CU_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    B_FAIL(B200_ECUDA,
           "b200: device %d is sm_%d%d; this library is built for sm_100a only",
           device, prop.major, prop.minor);
  b200_ctx *c = new b200_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  CU_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  CU_TRY(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  CU_TRY(cudaEventCreate(&c->ev_a));
  CU_TRY(cudaEventCreate(&c->ev_b));
  CU_TRY(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&c->ev_poll, cudaEventDisableTiming));
  CU_TRY(cudaHostAlloc((void **)&c->h_flag, 64, cudaHostAllocDefault));
  memset(c->h_flag, 0, 64);
  *out = c;
  return B200_OK;
}

extern "C" int b200_ctx_create(int device, b200_ctx **ctx) {
  return ctx_common(device, ctx);
}

int dist_comm_init(b200_ctx *c, const void *nccl_id);
void dist_comm_destroy(b200_ctx *c);

extern "C" int b200_ctx_create_dist(int device, int rank, int nranks,
                                    const void *nccl_id, b200_ctx **ctx) {
  if (nranks < 1 || rank < 0 || rank >= nranks)
    B_FAIL(B200_EINVAL, "b200_ctx_create_dist: rank %d of %d", rank, nranks);
  B_TRY(ctx_common(device, ctx));
  (*ctx)->rank = rank, (*ctx)->nranks = nranks;
  if (nranks > 1) {
    if (!nccl_id)
      B_FAIL(B200_EINVAL, "b200_ctx_create_dist: nccl_id is required");
    B_TRY(dist_comm_init(*ctx, nccl_id));
  }
  return B200_OK;
}

extern "C" int b200_ctx_destroy(b200_ctx *c) {
  if (!c)
    return B200_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  dist_comm_destroy(c);
  if (c->ev_a) cudaEventDestroy(c->ev_a);
  if (c->ev_b) cudaEventDestroy(c->ev_b);
  if (c->ev_halo) cudaEventDestroy(c->ev_halo);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->ev_poll) cudaEventDestroy(c->ev_poll);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->h_flag) cudaFreeHost(c->h_flag);
  delete c;
  return B200_OK;
}

extern "C" int b200_ctx_set_stream(b200_ctx *c, void *s) {
  if (!c)
    B_FAIL(B200_EINVAL, "b200_ctx_set_stream: null ctx");
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return B200_OK;
}

extern "C" int b200_ctx_sync(b200_ctx *c) {
  if (!c)
    B_FAIL(B200_EINVAL, "b200_ctx_sync: null ctx");
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaStreamSynchronize(c->stream));
  CU_TRY(cudaStreamSynchronize(c->comm_stream));
  return B200_OK;
}

extern "C" int b200_ctx_rank(const b200_ctx *c, int *rank, int *nranks) {
  if (!c)
    B_FAIL(B200_EINVAL, "b200_ctx_rank: null ctx");
  if (rank) *rank = c->rank;
  if (nranks) *nranks = c->nranks;
  return B200_OK;
}

extern "C" int b200_malloc(b200_ctx *c, size_t bytes, void **dptr) {
  if (!c || !dptr)
    B_FAIL(B200_EINVAL, "b200_malloc: null argument");
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaMalloc(dptr, bytes ? bytes : 8));
  return B200_OK;
}

extern "C" int b200_free(b200_ctx *c, void *dptr) {
  if (!c)
    B_FAIL(B200_EINVAL, "b200_free: null ctx");
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaFree(dptr));
  return B200_OK;
}

extern "C" int b200_memcpy_h2d(b200_ctx *c, void *dst, const void *src,
                               size_t bytes) {
  if (!c)
    B_FAIL(B200_EINVAL, "b200_memcpy_h2d: null ctx");
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return B200_OK;
}

extern "C" int b200_memcpy_d2h(b200_ctx *c, void *dst, const void *src,
                               size_t bytes) {
  if (!c)
    B_FAIL(B200_EINVAL, "b200_memcpy_d2h: null ctx");
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(cudaStreamSynchronize(c->stream));
  return B200_OK;
}

extern "C" int b200_memset(b200_ctx *c, void *dst, int byte, size_t bytes) {
  if (!c)
    B_FAIL(B200_EINVAL, "b200_memset: null ctx");
  CU_TRY(cudaSetDevice(c->device));
  CU_TRY(cudaMemsetAsync(dst, byte, bytes, c->stream));
  return B200_OK;
}

extern "C" int b200_host_alloc(size_t bytes, void **hptr) {
  if (!hptr)
    B_FAIL(B200_EINVAL, "b200_host_alloc: null argument");
  CU_TRY(cudaHostAlloc(hptr, bytes ? bytes : 8, cudaHostAllocDefault));
  return B200_OK;
}

extern "C" int b200_host_free(void *hptr) {
  CU_TRY(cudaFreeHost(hptr));
  return B200_OK;
}

int dev_alloc(b200_mat *M, void **p, size_t bytes) {
  *p = nullptr;
  CU_TRY(cudaMalloc(p, bytes ? bytes : 8));
  M->device_bytes += bytes;
  return B200_OK;
}
