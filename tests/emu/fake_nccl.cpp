/* TEST INFRASTRUCTURE ONLY -- an in-process stand-in for the few NCCL entry points
 * csrc/gfmd_b200.cu loads with dlopen, for the CUDA-on-CPU emulation build: the "ranks" of a
 * communicator are host threads of one process (tests/test_emulated_multi_rank.py), collectives
 * are real barriers between them, so the ordering the library relies on (every rank's pushes
 * precede its contribution to the all-reduce that serves as barrier) holds exactly.
 * Selected with GFMD_B200_NCCL_LIB=<this library>.  Doubles only. */
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "nccl.h"

typedef struct emu_stream *cudaStream_t;

namespace {

struct Op { bool send; void *buf; size_t bytes; int peer; bool used; };

struct Group {
  int n = 0, joined = 0;
  std::mutex m;
  std::condition_variable cv;
  int waiting = 0;
  unsigned gen = 0;
  std::vector<const void *> src;
  std::vector<std::vector<Op>> mail;
  void barrier()
  {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++waiting == n) {
      waiting = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};

std::mutex g_groups_mutex;
std::map<std::string, Group *> g_groups;
unsigned long long g_next_id = 1;

thread_local std::vector<Op> t_ops;
thread_local ncclComm_t t_comm = nullptr;
thread_local int t_depth = 0;

}  // namespace

struct ncclComm { Group *group; int rank; };

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId *id)
{
  std::lock_guard<std::mutex> lk(g_groups_mutex);
  memset(id, 0, sizeof(*id));
  const unsigned long long v = g_next_id++;
  memcpy(id->internal, &v, sizeof(v));
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank)
{
  Group *g;
  {
    std::lock_guard<std::mutex> lk(g_groups_mutex);
    const std::string key(id.internal, sizeof(id.internal));
    Group *&slot = g_groups[key];
    if (!slot) {
      slot = new Group;
      slot->n = nranks;
      slot->src.resize(nranks);
      slot->mail.resize(nranks);
    }
    g = slot;
  }
  *comm = new ncclComm{g, rank};
  g->barrier();                         /* like NCCL: returns when every rank has joined */
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm)
{
  delete comm;
  return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t, ncclRedOp_t, ncclComm_t comm,
                           cudaStream_t)
{
  Group *g = comm->group;
  g->src[comm->rank] = send;
  g->barrier();
  std::vector<double> sum(count, 0.0);
  for (int p = 0; p < g->n; ++p)
    for (size_t i = 0; i < count; ++i) sum[i] += static_cast<const double *>(g->src[p])[i];
  g->barrier();                         /* everybody has read every contribution */
  memcpy(recv, sum.data(), count * sizeof(double));
  return ncclSuccess;
}

ncclResult_t ncclGroupStart()
{
  ++t_depth;
  return ncclSuccess;
}

static ncclResult_t flush_ops()
{
  if (t_ops.empty() || !t_comm) return ncclSuccess;
  Group *g = t_comm->group;
  const int me = t_comm->rank;
  g->mail[me].clear();
  for (const Op &o : t_ops)
    if (o.send) g->mail[me].push_back(o);
  g->barrier();
  for (const Op &o : t_ops) {
    if (o.send) continue;
    bool found = false;
    for (Op &s : g->mail[o.peer])
      if (!found && s.peer == me && !s.used && s.bytes == o.bytes) {
        memcpy(o.buf, s.buf, o.bytes);
        s.used = true;                  /* only this rank consumes sends addressed to it */
        found = true;
      }
    if (!found) return ncclUnhandledCudaError;
  }
  g->barrier();
  t_ops.clear();
  return ncclSuccess;
}

ncclResult_t ncclGroupEnd()
{
  if (--t_depth > 0) return ncclSuccess;
  return flush_ops();
}

ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t, int peer, ncclComm_t comm, cudaStream_t)
{
  t_comm = comm;
  t_ops.push_back(Op{true, const_cast<void *>(buf), count * sizeof(double), peer, false});
  return t_depth > 0 ? ncclSuccess : flush_ops();
}

ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t, int peer, ncclComm_t comm, cudaStream_t)
{
  t_comm = comm;
  t_ops.push_back(Op{false, buf, count * sizeof(double), peer, false});
  return t_depth > 0 ? ncclSuccess : flush_ops();
}

const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "fake NCCL: unmatched send/recv"; }

}
