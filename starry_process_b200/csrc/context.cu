// Context management, error reporting, launch accounting for libspb200.
#include <mutex>

#include "common.cuh"
#include "spb_tables.h"

int spb_wigner_init(spb_context *ctx);  // wigner.cu: constant-bank copy of the Rx(pi/2) non-zeros

static thread_local std::string g_last_error;

void spb_set_error(const std::string &msg) { g_last_error = msg; }

extern "C" const char *spb_last_error(void) { return g_last_error.c_str(); }
extern "C" int spb_version(void) { return 100; }

extern "C" int spb_create(int device, const double *tables_host, size_t tables_count,
                          spb_context **out) {
  SPB_REQUIRE(out != nullptr, "spb_create: null output pointer");
  int ndev = 0;
  SPB_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  SPB_REQUIRE(device >= 0 && device < ndev, "spb_create: no such CUDA device");
  SPB_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SPB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  SPB_REQUIRE(prop.major >= 10, "spb_create: libspb200 is built for sm_100a (B200) only");
  spb_context *ctx = new spb_context();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->d_tables = nullptr;
  ctx->tables_count = tables_count;
  ctx->launches = 0;
  if (tables_count > 0) {
    SPB_REQUIRE(tables_host != nullptr, "spb_create: null table blob");
    SPB_CHECK_CUDA(cudaMalloc(&ctx->d_tables, tables_count * sizeof(double)));
    SPB_CHECK_CUDA(cudaMemcpy(ctx->d_tables, tables_host, tables_count * sizeof(double),
                              cudaMemcpyHostToDevice));
    if (tables_count == SPB_TAB_TOTAL) {
      int st = spb_wigner_init(ctx);
      if (st) return st;
    }
  }
  *out = ctx;
  return 0;
}

extern "C" void spb_destroy(spb_context *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  delete ctx;
}

extern "C" int spb_device(const spb_context *ctx) { return ctx ? ctx->device : -1; }

extern "C" int spb_launch_count(const spb_context *ctx, long long *count_host) {
  SPB_REQUIRE(ctx != nullptr && count_host != nullptr, "spb_launch_count: null argument");
  *count_host = ctx->launches;
  return 0;
}
