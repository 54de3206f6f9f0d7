// Multi-GPU plumbing inside the library: one process per GPU, an NCCL communicator owned by the context.
//
// MSMs and multi-Miller loops shard by input slices; the tiny per-GPU partials (one Fq12 Miller
// value or one group element per shard) are combined with ONE ncclAllGather over NVLink and multiplied / added on every
// rank in rank order (GT products and EC sums are not NCCL reduction operators).  The reference has no counterpart
// (rayon on one host); this is what a Rust host behind the traits calls to span the 8 GPUs of a box.
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy the process already holds when PyTorch was imported
// first, the system one otherwise), so the library has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>

#include "common.cuh"

namespace {
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
typedef int NcclResult;
struct NcclApi {
  void* h;
  NcclResult (*GetUniqueId)(NcclUniqueId*);
  NcclResult (*CommInitRank)(NcclComm*, int, NcclUniqueId, int);
  NcclResult (*CommDestroy)(NcclComm);
  NcclResult (*AllGather)(const void*, void*, size_t, int /*ncclDataType_t*/, NcclComm, cudaStream_t);
  const char* (*GetErrorString)(NcclResult);
  const char* (*GetLastError)(NcclComm);
};
const int NCCL_INT8 = 0;  // ncclInt8 = ncclChar = 0 (nccl.h)

NcclApi* nccl() {
  static NcclApi api = [] {
    NcclApi a;
    memset(&a, 0, sizeof(a));
    const char* names[] = {getenv("RIPP_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.h) break;
    }
    if (!a.h) return a;
    a.GetUniqueId = (NcclResult(*)(NcclUniqueId*))dlsym(a.h, "ncclGetUniqueId");
    a.CommInitRank = (NcclResult(*)(NcclComm*, int, NcclUniqueId, int))dlsym(a.h, "ncclCommInitRank");
    a.CommDestroy = (NcclResult(*)(NcclComm))dlsym(a.h, "ncclCommDestroy");
    a.AllGather = (NcclResult(*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(a.h, "ncclAllGather");
    a.GetErrorString = (const char* (*)(NcclResult))dlsym(a.h, "ncclGetErrorString");
    a.GetLastError = (const char* (*)(NcclComm))dlsym(a.h, "ncclGetLastError");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString) a.h = nullptr;
    return a;
  }();
  return api.h ? &api : nullptr;
}
int nccl_fail(const char* what, NcclResult r) {
  NcclApi* n = nccl();
  return fail(RIPP_ERR_NCCL, std::string(what) + ": " + (n ? n->GetErrorString(r) : "NCCL not loaded"));
}
}  // namespace

extern "C" int ripp_comm_unique_id(uint8_t* id_out) {
  if (!id_out) return fail(RIPP_ERR_ARG, "null argument");
  NcclApi* n = nccl();
  if (!n) return fail(RIPP_ERR_NCCL, std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "symbols missing"));
  NcclUniqueId id;
  NcclResult r = n->GetUniqueId(&id);
  if (r != 0) return nccl_fail("ncclGetUniqueId", r);
  memcpy(id_out, id.internal, 128);
  return RIPP_OK;
}

extern "C" int ripp_comm_init(ripp_ctx* ctx, const uint8_t* id, int rank, int world) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return fail(RIPP_ERR_ARG, "bad communicator arguments");
  if (ctx->parent) return fail(RIPP_ERR_ARG, "communicators belong to top-level contexts");
  if (world & (world - 1)) return fail(RIPP_ERR_ARG, "world size must be a power of two (cyclic partition of GIPA, SURVEY.md §8e)");
  ripp_comm_release(ctx);
  ctx->rank = rank;
  ctx->world = world;
  if (world == 1) return RIPP_OK;
  NcclApi* n = nccl();
  if (!n) return fail(RIPP_ERR_NCCL, "cannot load NCCL (libnccl.so.2)");
  CU(cudaSetDevice(ctx->device));
  NcclUniqueId uid;
  memcpy(uid.internal, id, 128);
  NcclComm c = nullptr;
  NcclResult r = n->CommInitRank(&c, world, uid, rank);
  if (r != 0) {
    ctx->world = 1;
    ctx->rank = 0;
    return nccl_fail("ncclCommInitRank", r);
  }
  ctx->comm = c;
  return RIPP_OK;
}

extern "C" int ripp_comm_info(ripp_ctx* ctx, int* rank, int* world) {
  if (!ctx) return fail(RIPP_ERR_ARG, "null ctx");
  if (rank) *rank = ctx->rank;
  if (world) *world = ctx->world;
  return RIPP_OK;
}

void ripp_comm_release(ripp_ctx* ctx) {
  if (ctx->comm) {
    NcclApi* n = nccl();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (n) n->CommDestroy((NcclComm)ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->rank = 0;
  ctx->world = 1;
}
extern "C" int ripp_comm_destroy(ripp_ctx* ctx) {
  if (!ctx) return fail(RIPP_ERR_ARG, "null ctx");
  ripp_comm_release(ctx);
  return RIPP_OK;
}

int ripp_all_gather_internal(ripp_ctx* ctx, const void* send_dev, size_t bytes, void* recv_dev) {
  ripp_ctx* top = ctx;
  while (top->parent) top = top->parent;
  if (top->world == 1) {
    if (send_dev != recv_dev) CU(cudaMemcpyAsync(recv_dev, send_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return RIPP_OK;
  }
  NcclApi* n = nccl();
  if (!n || !top->comm) return fail(RIPP_ERR_NCCL, "no communicator: call ripp_comm_init first");
  NcclResult r = n->AllGather(send_dev, recv_dev, bytes, NCCL_INT8, (NcclComm)top->comm, ctx->stream);
  if (r != 0) return nccl_fail("ncclAllGather", r);
  return RIPP_OK;
}

extern "C" int ripp_all_gather_dev(ripp_ctx* ctx, const void* send_dev, size_t bytes, void* recv_dev) {
  if (!ctx || !send_dev || !recv_dev) return fail(RIPP_ERR_ARG, "null argument");
  CU(cudaSetDevice(ctx->device));
  return ripp_all_gather_internal(ctx, send_dev, bytes, recv_dev);
}
