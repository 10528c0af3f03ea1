// Host stand-in for NCCL used ONLY by the emulated-device build (tools/emu): ranks are THREADS of one process, a
// communicator is a shared mailbox keyed by the unique id, sends are buffered and receives wait for them.  It lets the
// partition / offset / AllToAllv logic of blaze_b200/csrc/exchange.cu run on a CPU-only box; it is never part of the product.
#pragma once
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <vector>

#include "cuda_runtime.h"

typedef enum { ncclSuccess = 0, ncclInternalError = 3 } ncclResult_t;
typedef enum { ncclUint8 = 1, ncclUint64 = 5 } ncclDataType_t;
typedef struct { char internal[128]; } ncclUniqueId;

namespace emu_nccl {
struct World {
  std::mutex mu; std::condition_variable cv;
  std::map<std::tuple<int, int, long long>, std::vector<unsigned char>> box;      // (src, dst, seq) -> payload
  std::map<std::pair<int, int>, long long> send_seq, recv_seq;
};
struct Op { bool send; const void* sp; void* rp; size_t bytes; int peer; };
struct Comm { std::shared_ptr<World> w; int rank, world; };
inline std::mutex g_mu;
inline std::map<long long, std::shared_ptr<World>> g_worlds;
inline long long g_next_id = 1;
inline thread_local int group_depth = 0;
inline thread_local std::vector<std::pair<Comm*, Op>> pending;
inline size_t dsize(ncclDataType_t t) { return t == ncclUint64 ? 8 : 1; }
inline void run(Comm* c, const Op& op) {
  World& w = *c->w;
  std::unique_lock<std::mutex> l(w.mu);
  if (op.send) {
    const long long q = w.send_seq[{c->rank, op.peer}]++;
    w.box[{c->rank, op.peer, q}] = std::vector<unsigned char>((const unsigned char*)op.sp, (const unsigned char*)op.sp + op.bytes);
    w.cv.notify_all();
  } else {
    const long long q = w.recv_seq[{op.peer, c->rank}]++;
    const auto key = std::make_tuple(op.peer, c->rank, q);
    w.cv.wait(l, [&] { return w.box.count(key) != 0; });
    auto it = w.box.find(key);
    if (it->second.size() != op.bytes) abort();                     // a count mismatch is a bug in the caller
    memcpy(op.rp, it->second.data(), op.bytes); w.box.erase(it);
  }
}
inline void flush() { auto ops = std::move(pending); pending.clear(); for (auto& o : ops) if (o.second.send) run(o.first, o.second); for (auto& o : ops) if (!o.second.send) run(o.first, o.second); }
}  // namespace emu_nccl
typedef emu_nccl::Comm* ncclComm_t;

inline ncclResult_t ncclGetUniqueId(ncclUniqueId* id) { std::lock_guard<std::mutex> l(emu_nccl::g_mu); memset(id, 0, sizeof(*id)); const long long v = emu_nccl::g_next_id++; memcpy(id->internal, &v, 8); return ncclSuccess; }
inline ncclResult_t ncclCommInitRank(ncclComm_t* c, int world, ncclUniqueId id, int rank) {
  long long v; memcpy(&v, id.internal, 8);
  std::lock_guard<std::mutex> l(emu_nccl::g_mu);
  auto& w = emu_nccl::g_worlds[v]; if (!w) w = std::make_shared<emu_nccl::World>();
  *c = new emu_nccl::Comm{w, rank, world}; return ncclSuccess;
}
inline ncclResult_t ncclCommDestroy(ncclComm_t c) { delete c; return ncclSuccess; }
inline ncclResult_t ncclGroupStart() { emu_nccl::group_depth++; return ncclSuccess; }
inline ncclResult_t ncclGroupEnd() { if (--emu_nccl::group_depth == 0) emu_nccl::flush(); return ncclSuccess; }
inline ncclResult_t ncclSend(const void* p, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  emu_nccl::pending.push_back({c, {true, p, nullptr, n * emu_nccl::dsize(t), peer}}); if (!emu_nccl::group_depth) emu_nccl::flush(); return ncclSuccess; }
inline ncclResult_t ncclRecv(void* p, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  emu_nccl::pending.push_back({c, {false, nullptr, p, n * emu_nccl::dsize(t), peer}}); if (!emu_nccl::group_depth) emu_nccl::flush(); return ncclSuccess; }
inline ncclResult_t ncclAllGather(const void* sp, void* rp, size_t count, ncclDataType_t t, ncclComm_t c, cudaStream_t s) {
  ncclGroupStart();
  for (int p = 0; p < c->world; p++) { ncclSend(sp, count, t, p, c, s); ncclRecv((char*)rp + (size_t)p * count * emu_nccl::dsize(t), count, t, p, c, s); }
  return ncclGroupEnd();
}
inline const char* ncclGetErrorString(ncclResult_t) { return "emulated NCCL error"; }
