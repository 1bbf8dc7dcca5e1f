// comm.cu - the exchange step of the path behind the C ABI: NCCL sum-reduce of the accumulation planes over NVLink / NVSwitch.
//
// Replaces device_handle_result_sharing -> device_result_interface.c:107-299 of the reference (D2H into pinned host memory, H2D on
// the main device, `buffer_add`; at most 4 devices) and round 1's serial cudaMemcpyPeer + add loop of the C host. Every
// (pixel, sample id) pair is independent, so the only collective of the path is ONE in-place ncclReduce of the packed planes
// [sum R | sum G | sum B | sum luminance(colour^2)] (4 * width * height floats) onto the device that resolves the output, queued on
// the devices' own render streams (no host synchronisation: the reduce starts when the last sample pass of each device retires).
// The adaptive sampler's stage counts travel the other way with one ncclBroadcast of the per-block words.
//
// Two ways to form a communicator, both without any dependency on a launcher:
//   * lumb200_comm_create_all   one process drives several devices (Luminary's device manager; `LuminaryB200 --device 0xFF`):
//                               ncclCommInitAll, collectives issued for all devices inside one ncclGroup;
//   * lumb200_comm_create_rank  one process per device (bench.py under torchrun): rank 0 makes a 128-byte id
//                               (lumb200_comm_get_unique_id), the application hands it to the other ranks by whatever means it has.
// NCCL is bound at run time (dlopen of libnccl.so.2, the system's 2.27 or the copy a host process has already loaded), so
// liblumb200.so itself carries no link-time dependency on it: single-GPU users never load NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "lumb200_internal.cuh"

extern "C" Lumb200Result lumb200_device_get_cuda_index(Lumb200Device* device, uint32_t* index);

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                                                          = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                                                    = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*)                                                            = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)                                                                              = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t)       = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)         = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)                 = nullptr;
  ncclResult_t (*GroupStart)()                                                                                         = nullptr;
  ncclResult_t (*GroupEnd)()                                                                                           = nullptr;
  const char* (*GetErrorString)(ncclResult_t)                                                                          = nullptr;
  ncclResult_t (*GetVersion)(int*)                                                                                     = nullptr;
};

NcclApi g_nccl;

template <typename F>
bool bind(F& fn, const char* name) {
  fn = (F) dlsym(g_nccl.handle, name);
  return fn != nullptr;
}

Lumb200Result load_nccl() {
  if (g_nccl.handle)
    return LUMB200_SUCCESS;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle)
      break;
  }
  if (!g_nccl.handle) {
    lumb200_set_last_error("NCCL is not available (dlopen libnccl.so.2: %s)", dlerror());
    return LUMB200_ERROR_MISSING_DATA;
  }
  const bool ok = bind(g_nccl.GetUniqueId, "ncclGetUniqueId") && bind(g_nccl.CommInitRank, "ncclCommInitRank") &&
                  bind(g_nccl.CommInitAll, "ncclCommInitAll") && bind(g_nccl.CommDestroy, "ncclCommDestroy") && bind(g_nccl.Reduce, "ncclReduce") &&
                  bind(g_nccl.AllReduce, "ncclAllReduce") && bind(g_nccl.Broadcast, "ncclBroadcast") && bind(g_nccl.GroupStart, "ncclGroupStart") &&
                  bind(g_nccl.GroupEnd, "ncclGroupEnd") && bind(g_nccl.GetErrorString, "ncclGetErrorString") && bind(g_nccl.GetVersion, "ncclGetVersion");
  if (!ok) {
    lumb200_set_last_error("libnccl does not export the expected entry points");
    dlclose(g_nccl.handle);
    g_nccl.handle = nullptr;
    return LUMB200_ERROR_MISSING_DATA;
  }
  return LUMB200_SUCCESS;
}

}  // namespace

#define NCCL_TRY(expr)                                                                                         \
  do {                                                                                                         \
    ncclResult_t _r = (expr);                                                                                  \
    if (_r != ncclSuccess) {                                                                                   \
      lumb200_set_last_error("NCCL error %d (%s) in %s at %s:%d", (int) _r, g_nccl.GetErrorString(_r), #expr, __FILE__, __LINE__); \
      return LUMB200_ERROR_CUDA;                                                                               \
    }                                                                                                          \
  } while (0)

#define CM_TRY(expr)                 \
  do {                               \
    Lumb200Result _r = (expr);       \
    if (_r != LUMB200_SUCCESS)       \
      return _r;                     \
  } while (0)

struct Lumb200Comm {
  ncclComm_t comm       = nullptr;
  Lumb200Device* device = nullptr;
  uint32_t world        = 1;
  uint32_t rank         = 0;
  int cuda_index        = 0;
};

static_assert(sizeof(ncclUniqueId) <= LUMB200_COMM_ID_BYTES, "ncclUniqueId must fit the id blob of the C ABI");

extern "C" Lumb200Result lumb200_comm_get_unique_id(void* id) {
  if (!id) {
    lumb200_set_last_error("id is NULL");
    return LUMB200_ERROR_ARGUMENT_NULL;
  }
  CM_TRY(load_nccl());
  ncclUniqueId uid;
  NCCL_TRY(g_nccl.GetUniqueId(&uid));
  memset(id, 0, LUMB200_COMM_ID_BYTES);
  memcpy(id, &uid, sizeof(uid));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_comm_create_rank(Lumb200Comm** comm, Lumb200Device* device, uint32_t world_size, uint32_t rank, const void* id) {
  if (!comm || !device || !id) {
    lumb200_set_last_error("NULL argument");
    return LUMB200_ERROR_ARGUMENT_NULL;
  }
  *comm = nullptr;
  if (world_size == 0 || rank >= world_size) {
    lumb200_set_last_error("rank %u is outside a communicator of %u ranks", rank, world_size);
    return LUMB200_ERROR_INVALID_API_ARGUMENT;
  }
  CM_TRY(load_nccl());
  uint32_t index = 0;
  CM_TRY(lumb200_device_get_cuda_index(device, &index));
  LB_CHECK(cudaSetDevice((int) index));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  Lumb200Comm* c = new Lumb200Comm();
  c->device      = device;
  c->world       = world_size;
  c->rank        = rank;
  c->cuda_index  = (int) index;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, (int) world_size, uid, (int) rank);
  if (r != ncclSuccess) {
    lumb200_set_last_error("ncclCommInitRank(%u of %u) failed: %s", rank, world_size, g_nccl.GetErrorString(r));
    delete c;
    return LUMB200_ERROR_CUDA;
  }
  *comm = c;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_comm_create_all(Lumb200Comm** comms, Lumb200Device* const* devices, uint32_t count) {
  if (!comms || !devices) {
    lumb200_set_last_error("NULL argument");
    return LUMB200_ERROR_ARGUMENT_NULL;
  }
  if (count == 0 || count > 64) {
    lumb200_set_last_error("a communicator needs 1..64 devices, got %u", count);
    return LUMB200_ERROR_INVALID_API_ARGUMENT;
  }
  for (uint32_t k = 0; k < count; k++)
    comms[k] = nullptr;
  CM_TRY(load_nccl());
  std::vector<int> indices(count);
  for (uint32_t k = 0; k < count; k++) {
    if (!devices[k]) {
      lumb200_set_last_error("device %u is NULL", k);
      return LUMB200_ERROR_ARGUMENT_NULL;
    }
    uint32_t index = 0;
    CM_TRY(lumb200_device_get_cuda_index(devices[k], &index));
    indices[k] = (int) index;
    for (uint32_t j = 0; j < k; j++)
      if (indices[j] == indices[k]) {
        lumb200_set_last_error("CUDA device %d appears twice in the communicator", indices[k]);
        return LUMB200_ERROR_INVALID_API_ARGUMENT;
      }
  }
  std::vector<ncclComm_t> raw(count, nullptr);
  NCCL_TRY(g_nccl.CommInitAll(raw.data(), (int) count, indices.data()));
  for (uint32_t k = 0; k < count; k++) {
    Lumb200Comm* c = new Lumb200Comm();
    c->comm        = raw[k];
    c->device      = devices[k];
    c->world       = count;
    c->rank        = k;
    c->cuda_index  = indices[k];
    comms[k]       = c;
  }
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_comm_destroy(Lumb200Comm** comm) {
  if (!comm) {
    lumb200_set_last_error("comm is NULL");
    return LUMB200_ERROR_ARGUMENT_NULL;
  }
  Lumb200Comm* c = *comm;
  if (!c)
    return LUMB200_SUCCESS;
  if (c->comm && g_nccl.CommDestroy) {
    cudaSetDevice(c->cuda_index);
    g_nccl.CommDestroy(c->comm);
  }
  delete c;
  *comm = nullptr;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_comm_get_info(Lumb200Comm* comm, uint32_t* world_size, uint32_t* rank, uint32_t* nccl_version) {
  if (!comm) {
    lumb200_set_last_error("comm is NULL");
    return LUMB200_ERROR_ARGUMENT_NULL;
  }
  if (world_size)
    *world_size = comm->world;
  if (rank)
    *rank = comm->rank;
  if (nccl_version) {
    int v = 0;
    NCCL_TRY(g_nccl.GetVersion(&v));
    *nccl_version = (uint32_t) v;
  }
  return LUMB200_SUCCESS;
}

// queues the in-place plane reduce of ONE member; the caller brackets several members of one process with a group
static Lumb200Result queue_reduce(Lumb200Comm* c, uint32_t root) {
  void* planes = nullptr;
  size_t n     = 0;
  void* stream = nullptr;
  CM_TRY(lumb200_device_get_frame_planes(c->device, &planes, &n));
  CM_TRY(lumb200_device_get_stream(c->device, &stream));
  LB_CHECK(cudaSetDevice(c->cuda_index));
  NCCL_TRY(g_nccl.Reduce(planes, planes, n, ncclFloat, ncclSum, (int) root, c->comm, (cudaStream_t) stream));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_comm_reduce_planes(Lumb200Comm* comm, uint32_t root) {
  if (!comm) {
    lumb200_set_last_error("comm is NULL");
    return LUMB200_ERROR_ARGUMENT_NULL;
  }
  if (root >= comm->world) {
    lumb200_set_last_error("root %u is outside the communicator (%u ranks)", root, comm->world);
    return LUMB200_ERROR_INVALID_API_ARGUMENT;
  }
  if (comm->world == 1)
    return LUMB200_SUCCESS;
  return queue_reduce(comm, root);
}

extern "C" Lumb200Result lumb200_comm_reduce_planes_all(Lumb200Comm* const* comms, uint32_t count, uint32_t root) {
  if (!comms) {
    lumb200_set_last_error("comms is NULL");
    return LUMB200_ERROR_ARGUMENT_NULL;
  }
  if (count <= 1)
    return LUMB200_SUCCESS;
  for (uint32_t k = 0; k < count; k++)
    if (!comms[k] || comms[k]->world != count || root >= count) {
      lumb200_set_last_error("the communicator list must hold all %u members of one communicator and a valid root", count);
      return LUMB200_ERROR_INVALID_API_ARGUMENT;
    }
  NCCL_TRY(g_nccl.GroupStart());
  Lumb200Result r = LUMB200_SUCCESS;
  for (uint32_t k = 0; k < count && r == LUMB200_SUCCESS; k++)
    r = queue_reduce(comms[k], root);
  NCCL_TRY(g_nccl.GroupEnd());
  return r;
}

// adaptive sampler: the stage counts rank `root` built travel to every member (one 32-bit word per 4 x 4 block)
static Lumb200Result queue_broadcast_words(Lumb200Comm* c, uint32_t root) {
  void *words = nullptr, *prefix = nullptr;
  size_t n     = 0;
  void* stream = nullptr;
  CM_TRY(lumb200_device_get_adaptive_words_device(c->device, &words, &prefix, &n));
  CM_TRY(lumb200_device_get_stream(c->device, &stream));
  LB_CHECK(cudaSetDevice(c->cuda_index));
  // the stage counts and the inclusive prefix sums of the tasks per block that the root derived from them
  NCCL_TRY(g_nccl.Broadcast(words, words, n, ncclUint32, (int) root, c->comm, (cudaStream_t) stream));
  NCCL_TRY(g_nccl.Broadcast(prefix, prefix, n, ncclUint32, (int) root, c->comm, (cudaStream_t) stream));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_comm_broadcast_adaptive_words(Lumb200Comm* comm, uint32_t root) {
  if (!comm) {
    lumb200_set_last_error("comm is NULL");
    return LUMB200_ERROR_ARGUMENT_NULL;
  }
  if (comm->world == 1)
    return LUMB200_SUCCESS;
  NCCL_TRY(g_nccl.GroupStart());
  const Lumb200Result r = queue_broadcast_words(comm, root);
  NCCL_TRY(g_nccl.GroupEnd());
  return r;
}

extern "C" Lumb200Result lumb200_comm_broadcast_adaptive_words_all(Lumb200Comm* const* comms, uint32_t count, uint32_t root) {
  if (!comms) {
    lumb200_set_last_error("comms is NULL");
    return LUMB200_ERROR_ARGUMENT_NULL;
  }
  if (count <= 1)
    return LUMB200_SUCCESS;
  NCCL_TRY(g_nccl.GroupStart());
  Lumb200Result r = LUMB200_SUCCESS;
  for (uint32_t k = 0; k < count && r == LUMB200_SUCCESS; k++)
    r = queue_broadcast_words(comms[k], root);
  NCCL_TRY(g_nccl.GroupEnd());
  return r;
}
