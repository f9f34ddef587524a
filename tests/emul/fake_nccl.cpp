// TEST INFRASTRUCTURE -- never part of the product, never loaded by it.
//
// Stand-in for the slice of NCCL the engine binds (user-eph_b200/csrc/eph_nccl.h), linked into the host build of the
// device sources so that the multi-rank data plane of the C ABI (eph_b200_comm_init, set_ghost_map, exchange_ghosts,
// reduce_and_solve) runs on CPU ranks.  "Device" memory is host memory here; the bytes move through callbacks the test
// registers (torch.distributed over gloo in tests/test_multirank_emulated.py).  Point-to-point operations between
// ncclGroupStart and ncclGroupEnd are handed over as one batch, in posting order, like NCCL matches them per peer.
#include <cstddef>
#include <cstring>
#include <vector>

#include "cuda_runtime.h"

namespace ephb {
struct NcclUniqueId { char internal[128]; };
struct ncclComm { int rank, size; };
typedef ncclComm *NcclComm;
}  // namespace ephb
using namespace ephb;

extern "C" {

typedef void (*emul_nccl_p2p_fn)(int nops, const int *is_send, const int *peer, void *const *buf, const size_t *bytes);
typedef void (*emul_nccl_allreduce_fn)(void *buf, size_t count);   // sum of doubles over all ranks, in place
typedef void (*emul_nccl_allgather_fn)(const void *send, void *recv, size_t bytes_per_rank);

static emul_nccl_p2p_fn g_p2p = nullptr;
static emul_nccl_allreduce_fn g_allreduce = nullptr;
static emul_nccl_allgather_fn g_allgather = nullptr;
static int g_depth = 0;
static std::vector<int> g_is_send, g_peer;
static std::vector<void *> g_buf;
static std::vector<size_t> g_bytes;
static long long g_calls[4] = {0, 0, 0, 0};   // groups flushed, p2p operations, all-reduces, all-gathers

void emul_nccl_set_transport(emul_nccl_p2p_fn p2p, emul_nccl_allreduce_fn ar, emul_nccl_allgather_fn ag) {
  g_p2p = p2p; g_allreduce = ar; g_allgather = ag;
}
long long emul_nccl_calls(int which) { return which >= 0 && which < 4 ? g_calls[which] : -1; }

static size_t type_bytes(int t) { return t == 8 ? 8 : (t == 7 || t == 2 || t == 3) ? 4 : (t == 4 || t == 5) ? 8 : 1; }

static int flush() {
  if (g_is_send.empty()) return 0;
  if (!g_p2p) return 1;
  ++g_calls[0];
  g_calls[1] += (long long)g_is_send.size();
  g_p2p((int)g_is_send.size(), g_is_send.data(), g_peer.data(), g_buf.data(), g_bytes.data());
  g_is_send.clear(); g_peer.clear(); g_buf.clear(); g_bytes.clear();
  return 0;
}

int ncclGetUniqueId(NcclUniqueId *id) { std::memset(id, 0, sizeof *id); std::strcpy(id->internal, "emulated"); return 0; }
int ncclCommInitRank(NcclComm *comm, int nranks, NcclUniqueId id, int rank) {
  if (std::strcmp(id.internal, "emulated") != 0) return 4;   // the id must have travelled from rank 0
  *comm = new ncclComm{rank, nranks};
  return 0;
}
int ncclCommDestroy(NcclComm comm) { delete comm; return 0; }
int ncclGroupStart() { ++g_depth; return 0; }
int ncclGroupEnd() { return --g_depth == 0 ? flush() : 0; }
static int post(int is_send, void *buf, size_t count, int type, int peer, NcclComm comm) {
  if (!comm || peer < 0 || peer >= comm->size || peer == comm->rank) return 4;
  g_is_send.push_back(is_send); g_peer.push_back(peer); g_buf.push_back(buf); g_bytes.push_back(count * type_bytes(type));
  return g_depth == 0 ? flush() : 0;
}
int ncclSend(const void *buf, size_t count, int type, int peer, NcclComm comm, cudaStream_t) { return post(1, const_cast<void *>(buf), count, type, peer, comm); }
int ncclRecv(void *buf, size_t count, int type, int peer, NcclComm comm, cudaStream_t) { return post(0, buf, count, type, peer, comm); }
int ncclAllReduce(const void *send, void *recv, size_t count, int type, int op, NcclComm comm, cudaStream_t) {
  if (!comm || type != 8 || op != 0) return 4;
  if (send != recv) std::memmove(recv, send, count * 8);
  ++g_calls[2];
  if (comm->size == 1) return 0;
  if (!g_allreduce) return 1;
  g_allreduce(recv, count);
  return 0;
}
int ncclAllGather(const void *send, void *recv, size_t count, int type, NcclComm comm, cudaStream_t) {
  if (!comm) return 4;
  ++g_calls[3];
  const size_t bytes = count * type_bytes(type);
  if (comm->size == 1) { if (send != recv) std::memmove(recv, send, bytes); return 0; }
  if (!g_allgather) return 1;
  g_allgather(send, recv, bytes);
  return 0;
}
// sum over ranks, rank r keeps block r: an all-reduce of a scratch copy through the same transport, then the block
int ncclReduceScatter(const void *send, void *recv, size_t count, int type, int op, NcclComm comm, cudaStream_t) {
  if (!comm || type != 8 || op != 0) return 4;
  ++g_calls[2];
  std::vector<double> tmp(static_cast<const double *>(send), static_cast<const double *>(send) + count * comm->size);
  if (comm->size > 1) {
    if (!g_allreduce) return 1;
    g_allreduce(tmp.data(), tmp.size());
  }
  std::memmove(recv, tmp.data() + count * comm->rank, count * 8);
  return 0;
}
const char *ncclGetErrorString(int r) { return r == 0 ? "no error" : r == 1 ? "emulated NCCL: no transport registered" : "emulated NCCL: invalid argument"; }

}  // extern "C"
