#include "djets_nccl.h"

#include <dlfcn.h>

#include <mutex>
#include <string>

namespace djets {

static NcclApi g_api;
static bool g_ok = false;
static std::string g_why;
static std::once_flag g_once;

static void bind_all() {
  // RTLD_NOLOAD first: reuse an NCCL the host process (PyTorch, Julia's NCCL_jll ...) already mapped.
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    const char* env = getenv("DJETS_NCCL_LIB");
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  }
  for (int i = 0; !h && i < 2; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    g_why = std::string("cannot dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "unknown");
    return;
  }
#define DJ_BIND(field, sym)                                          \
  g_api.field = reinterpret_cast<decltype(g_api.field)>(dlsym(h, sym)); \
  if (!g_api.field) {                                                \
    g_why = std::string("libnccl lacks ") + sym;                     \
    return;                                                          \
  }
  DJ_BIND(GetUniqueId, "ncclGetUniqueId");
  DJ_BIND(CommInitRank, "ncclCommInitRank");
  DJ_BIND(CommDestroy, "ncclCommDestroy");
  DJ_BIND(GroupStart, "ncclGroupStart");
  DJ_BIND(GroupEnd, "ncclGroupEnd");
  DJ_BIND(Send, "ncclSend");
  DJ_BIND(Recv, "ncclRecv");
  DJ_BIND(AllReduce, "ncclAllReduce");
  DJ_BIND(AllGather, "ncclAllGather");
  DJ_BIND(GetErrorString, "ncclGetErrorString");
  DJ_BIND(GetVersion, "ncclGetVersion");
#undef DJ_BIND
  g_ok = true;
}

const NcclApi* nccl_api(const char** why) {
  std::call_once(g_once, bind_all);
  if (!g_ok) {
    if (why) *why = g_why.c_str();
    return nullptr;
  }
  return &g_api;
}

}  // namespace djets
