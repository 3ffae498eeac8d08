// NVLink peer-memory exchange of the loss sums between the ranks of one node (device side), shared by the mining
// kernel (post / collect), the stand-alone collect kernel and the Detect tail that completes a deferred call.
#pragma once
#include "common.h"
#include "ssdbox_dev.cuh"

namespace ssdbox {

// ---- NVLink peer-memory reduction of {sum smooth-L1, sum CE, N} (see ssdbox_peer_group) ----------
// Exchange buffer of a rank: [0] call epoch (u64, touched by the owner only), [1] timeouts seen by the owner, then
// at byte 64 two banks (epoch parity) of `world` 64-byte slots.  Slot (parity, r) of rank q's buffer is written by
// rank r only.  A rank can run at most one call ahead of a peer's reads (it needs that peer's slot of the current
// call to finish), so two banks are enough.
// A slot carries the three fp64 sums as SIX self-validating 8-byte words { epoch tag : 32 | half of a double : 32 }
// (the scheme of NCCL's LL protocol): an aligned 8-byte store is single-copy atomic, so a reader that sees the tag
// of the current call in a word also sees that word's payload.  The poster therefore needs no release fence and no
// store round trip over NVLink -- six fire-and-forget stores per peer -- and the collector no acquire.
constexpr int kPeerHeaderBytes = 64;
constexpr int kPeerSlotBytes = 64;
constexpr int kPeerWords = 6;

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t peer_tag(unsigned long long epoch) { return (uint32_t)(epoch % 0xffffffffull) + 1u; }   // never 0

// one warp: this rank's sums -> slot `me` of every rank's buffer (peer stores over NVLink; own buffer
// for lane == me).  Returns the call epoch (advanced here, kept in the owner's buffer).
__device__ __forceinline__ unsigned long long peer_post(void* const* bufs, int me, int world, double v0, double v1, double v2, int lane) {
  unsigned long long* mine = reinterpret_cast<unsigned long long*>(bufs[me]);
  unsigned long long epoch = 0;
  if (lane == 0) {
    epoch = mine[0] + 1ull;
    mine[0] = epoch;
  }
  epoch = __shfl_sync(SSDBOX_FULL_MASK, epoch, 0);
  const size_t bank = kPeerHeaderBytes + (size_t)(epoch & 1ull) * world * kPeerSlotBytes;
  if (lane < world) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(static_cast<char*>(bufs[lane]) + bank + (size_t)me * kPeerSlotBytes);
    const unsigned long long tag = (unsigned long long)peer_tag(epoch) << 32;
    const unsigned long long b0 = (unsigned long long)__double_as_longlong(v0);
    const unsigned long long b1 = (unsigned long long)__double_as_longlong(v1);
    const unsigned long long b2 = (unsigned long long)__double_as_longlong(v2);
    st_relaxed_sys(dst + 0, tag | (b0 & 0xffffffffull));
    st_relaxed_sys(dst + 1, tag | (b0 >> 32));
    st_relaxed_sys(dst + 2, tag | (b1 & 0xffffffffull));
    st_relaxed_sys(dst + 3, tag | (b1 >> 32));
    st_relaxed_sys(dst + 4, tag | (b2 & 0xffffffffull));
    st_relaxed_sys(dst + 5, tag | (b2 >> 32));
  }
  return epoch;
}

// one warp: waits for every rank's slot of call `epoch` in MY buffer and adds them in rank order (the
// same fp64 result on every rank); result valid in lane 0
__device__ __forceinline__ unsigned long long wall_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// peers->wait_timeout_ms, 0 = default
static inline long long peer_timeout_ns(const ssdbox_peer_group* peers) {
  const long long ms = peers->wait_timeout_ms > 0 ? peers->wait_timeout_ms : 30000;
  return ms * 1000000ll;
}

__device__ __forceinline__ void peer_collect(void* const* bufs, int me, int world, unsigned long long epoch, int lane, double* out,
                             long long timeout_ns) {
  char* mine = static_cast<char*>(bufs[me]);
  const size_t bank = kPeerHeaderBytes + (size_t)(epoch & 1ull) * world * kPeerSlotBytes;
  double r0 = 0.0, r1 = 0.0, r2 = 0.0;
  bool arrived = true;
  if (lane < world) {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(mine + bank + (size_t)lane * kPeerSlotBytes);
    const uint32_t tag = peer_tag(epoch);
    unsigned long long w[kPeerWords];
    const unsigned long long t0 = wall_ns();
    unsigned spins = 0;
    for (;;) {
      bool all = true;
#pragma unroll
      for (int i = 0; i < kPeerWords; ++i) {
        w[i] = ld_relaxed_sys(src + i);
        all = all && (uint32_t)(w[i] >> 32) == tag;
      }
      if (all) break;
      if ((++spins & 1023u) == 0 && wall_ns() - t0 > (unsigned long long)timeout_ns) {     // a peer never arrived
        arrived = false;
        break;
      }
    }
    r0 = __longlong_as_double((long long)((w[0] & 0xffffffffull) | (w[1] << 32)));
    r1 = __longlong_as_double((long long)((w[2] & 0xffffffffull) | (w[3] << 32)));
    r2 = __longlong_as_double((long long)((w[4] & 0xffffffffull) | (w[5] << 32)));
  }
  // A missing peer does not kill the context: the sums of this call become NaN and the owner's buffer
  // counts the event (header word 1, read by PeerExchange.timeouts()).
  const bool ok = __all_sync(SSDBOX_FULL_MASK, arrived);
  double t0s = 0.0, t1s = 0.0, t2s = 0.0;
  for (int r = 0; r < world; ++r) {
    t0s += __shfl_sync(SSDBOX_FULL_MASK, r0, r);
    t1s += __shfl_sync(SSDBOX_FULL_MASK, r1, r);
    t2s += __shfl_sync(SSDBOX_FULL_MASK, r2, r);
  }
  if (!ok) {
    t0s = t1s = t2s = __longlong_as_double(0x7ff8000000000000ll);
    if (lane == 0) reinterpret_cast<unsigned long long*>(mine)[1] += 1ull;
  }
  out[0] = t0s;
  out[1] = t1s;
  out[2] = t2s;
}

struct PeerFinishArgs {
  int rank, world;
  long long timeout_ns;
  void* bufs[SSDBOX_MAX_PEERS];
  double* sums;
  float* losses;
};


// one warp: completes a call whose wait was deferred -- collect, write the global sums and (nullable) the losses
__device__ __forceinline__ void peer_finish_warp(const PeerFinishArgs& a, int lane) {
  const unsigned long long epoch = *reinterpret_cast<const unsigned long long*>(a.bufs[a.rank]);   // posted by the forward
  double g[3];
  peer_collect(a.bufs, a.rank, a.world, epoch, lane, g, a.timeout_ns);
  if (lane == 0) {
    a.sums[0] = g[0];
    a.sums[1] = g[1];
    a.sums[2] = g[2];
    if (a.losses) {
      a.losses[0] = g[2] == 0.0 ? 0.0f : (float)(g[0] / g[2]);
      a.losses[1] = g[2] == 0.0 ? 0.0f : (float)(g[1] / g[2]);
    }
  }
}

}  // namespace ssdbox
