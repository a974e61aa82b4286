// nccl_dl.cuh — the ONE collective of the path: the gradient all-reduce of tf.distribute.MirroredStrategy (train.py:203;
// implicit in optimizer.apply_gradients, model.py:336), as ncclAllReduce(SUM) over the flat fp32 gradient buffer.
//
// NCCL is resolved at run time (dlopen), not linked: a process that already carries a libnccl.so.2 (torch bundles one and
// loads it first) must end up with ONE copy of the library, and a single-GPU user needs none at all.  Only the stable C
// entry points of NCCL 2.x are used; the handful of types below mirror nccl.h (ncclUniqueId is 128 opaque bytes,
// ncclFloat32 = 7, ncclSum = 0 in every 2.x release).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

typedef struct ncclComm* wn_ncclComm_t;
typedef struct { char internal[128]; } wn_ncclUniqueId;
enum { WN_NCCL_FLOAT32 = 7, WN_NCCL_SUM = 0 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(wn_ncclUniqueId*) = nullptr;
  int (*CommInitRank)(wn_ncclComm_t*, int, wn_ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(wn_ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, wn_ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*GetVersion)(int*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  char where[256] = "";
};

static NcclApi g_nccl;

// 0 on success; the error text lands in `err`
static int nccl_load(char* err, size_t errlen) {
  if (g_nccl.lib) return 0;
  const char* env = getenv("WN_NCCL_LIB");
  void* lib = nullptr;
  if (env && env[0]) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  // a copy the process already carries (torch's bundled libnccl.so.2 registers under its SONAME) wins over the system one
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { snprintf(err, errlen, "cannot load libnccl.so.2 (%s); set WN_NCCL_LIB", dlerror()); return -1; }
  NcclApi a;
  a.lib = lib;
  a.GetUniqueId = (int (*)(wn_ncclUniqueId*))dlsym(lib, "ncclGetUniqueId");
  a.CommInitRank = (int (*)(wn_ncclComm_t*, int, wn_ncclUniqueId, int))dlsym(lib, "ncclCommInitRank");
  a.CommDestroy = (int (*)(wn_ncclComm_t))dlsym(lib, "ncclCommDestroy");
  a.AllReduce = (int (*)(const void*, void*, size_t, int, int, wn_ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllReduce");
  a.GroupStart = (int (*)())dlsym(lib, "ncclGroupStart");
  a.GroupEnd = (int (*)())dlsym(lib, "ncclGroupEnd");
  a.GetVersion = (int (*)(int*))dlsym(lib, "ncclGetVersion");
  a.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GroupStart || !a.GroupEnd) {
    snprintf(err, errlen, "libnccl.so.2 lacks a required entry point");
    return -2;
  }
  Dl_info info;
  if (dladdr((void*)a.AllReduce, &info) && info.dli_fname) snprintf(a.where, sizeof(a.where), "%s", info.dli_fname);
  g_nccl = a;
  return 0;
}
static inline const char* nccl_errstr(int r) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error"; }
