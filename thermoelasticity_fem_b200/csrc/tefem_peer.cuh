// One-shot all-reduce over NVLink / NVSwitch PEER MEMORY, fused into the kernels that produce and consume the
// reduced values (the BiCGSTAB dot-product packets and the restricted residual of the coarse correction).
//
// Every rank owns one cudaMalloc'ed region that all ranks of the node map with CUDA IPC:
//     [flags: channel x parity x source rank, uint32][small packets: parity x 16 doubles][big: parity x cap doubles]
// A reduction of epoch e (host counter per channel, identical on all ranks because every rank launches the same
// kernel sequence) uses the buffers of parity e & 1:
//   1. the rank's contribution is in ITS OWN buffer (written by this kernel or by the producer kernel before it);
//   2. it raises flag[channel][parity][my rank] = e in EVERY rank's region (st.release.sys after a system fence);
//   3. it waits until all its own flags of that channel / parity show e (ld.acquire.sys), then reads every rank's
//      buffer over NVLink (volatile loads: no stale L1 lines) and adds them IN RANK ORDER: every rank obtains the
//      bitwise identical sum, which keeps the replicated coarse levels and the stop decisions consistent.
// Reuse is safe with two parities: a rank reaches epoch e + 2 only after every peer signalled e + 1, i.e. after
// every peer finished reading the buffers of epoch e.  The waits are bounded (PEER_TIMEOUT cycles): on expiry the
// kernel records an error code that the host reports after its next synchronisation instead of hanging the GPU.
#pragma once
#include <stdint.h>

namespace tefem {

constexpr int PEER_MAX_RANKS = 8;
constexpr int PEER_SMALL = 16;                       // doubles per small packet
constexpr int64_t PEER_FLAG_BYTES = 1024;            // 4 channels x 2 parities x 8 ranks x 4 B, padded
constexpr int64_t PEER_SMALL_BYTES = 2 * PEER_SMALL * 8;
constexpr int64_t PEER_BIG_OFFSET = 2048;
constexpr long long PEER_TIMEOUT_DEFAULT = 60000000000LL;     // ~30 s at 1.9 GHz (tefem_comm_set_timeout overrides it)

// channels: dot-product packets, restricted residual of the coarse correction (replicated cycle), halo values of the SpMV
// input, and the application counter of the row-partitioned coarse cycle (its tags; that channel uses no flags)
enum { PEER_CH_DOTS = 0, PEER_CH_COARSE = 1, PEER_CH_HALO = 2, PEER_CH_CYCLE = 3, PEER_N_CHANNELS = 4 };

struct PeerCtx {
    int world = 1, rank = 0;
    uint32_t epoch = 0;
    int channel = 0;
    int64_t cap = 0;                                 // doubles per parity of the big buffer
    char* base[PEER_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int* err = nullptr;                              // device int: set to 1 on a wait time-out
    long long timeout = PEER_TIMEOUT_DEFAULT;        // cycles a wait may spin before it gives up
};

// flag[channel][parity][source rank] in `owner`'s region (host + device)
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t* peer_flag(const PeerCtx& c, int owner, int src) {
    return reinterpret_cast<uint32_t*>(c.base[owner]) + ((c.channel * 2 + (int)(c.epoch & 1u)) * PEER_MAX_RANKS + src);
}
#ifdef __CUDACC__
__device__ __forceinline__ double* peer_small(const PeerCtx& c, int owner) {
    return reinterpret_cast<double*>(c.base[owner] + PEER_FLAG_BYTES) + (c.epoch & 1u) * PEER_SMALL;
}
__device__ __forceinline__ double* peer_big(const PeerCtx& c, int owner) {
    return reinterpret_cast<double*>(c.base[owner] + PEER_BIG_OFFSET) + (int64_t)(c.epoch & 1u) * c.cap;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ld_volatile_v2(const double* p, double& a, double& b) {
    asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p) : "memory");
}

// Thread `t` (< world) raises this rank's flag in rank t's region.  Call after the contribution is globally
// visible (the caller issues __threadfence_system() + a CTA barrier, or the data comes from an earlier kernel).
__device__ __forceinline__ void peer_signal(const PeerCtx& c, int t) {
    if (t < c.world) st_release_sys(peer_flag(c, t, c.rank), c.epoch);
}
// Thread `t` (< world) waits for rank t's flag in the local region.  A wait that expires records the error and every
// later wait on this communicator returns at once (the host reports it after its next synchronisation and the
// communicator stays marked until tefem_comm_reset_error): a lost rank costs ONE time-out, not one per kernel.
__device__ __forceinline__ void peer_wait_flag(const PeerCtx& c, int src) {
    const uint32_t* f = peer_flag(c, c.rank, src);
    if ((int32_t)(ld_acquire_sys(f) - c.epoch) >= 0) return;
    if (*reinterpret_cast<volatile int*>(c.err)) return;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - c.epoch) < 0) {
        if (clock64() - t0 > c.timeout) { *c.err = 1; break; }
    }
}
__device__ __forceinline__ void peer_wait(const PeerCtx& c, int t) {
    if (t < c.world) peer_wait_flag(c, t);
}
#endif

}  // namespace tefem
