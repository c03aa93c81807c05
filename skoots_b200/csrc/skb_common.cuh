// Shared device/host helpers for the skoots_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "skoots_b200.h"

typedef unsigned long long ull;

void skb_set_error(const char* fmt, ...);

#define SKB_REQUIRE(cond, ...)        \
    do {                              \
        if (!(cond)) {                \
            skb_set_error(__VA_ARGS__); \
            return SKB_E_ARG;         \
        }                             \
    } while (0)

#define SKB_LAUNCH_CHECK(what)                                                   \
    do {                                                                         \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) {                                                \
            skb_set_error("%s: CUDA error %s", what, cudaGetErrorString(e__));   \
            return SKB_E_CUDA;                                                   \
        }                                                                        \
    } while (0)

// > 48 KB of dynamic shared memory is an opt-in per function AND per device: raise it the first time a kernel is launched
// on a device instead of on every launch (the call costs a few microseconds of host time on launch-bound paths).
#define SKB_RAISE_SMEM_ONCE(kernel, bytes)                                                                     \
    do {                                                                                                       \
        static bool raised__[64] = {};                                                                         \
        int dev__ = 0;                                                                                         \
        cudaGetDevice(&dev__);                                                                                 \
        if (dev__ < 0 || dev__ >= 64 || !raised__[dev__]) {                                                    \
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));           \
            if (dev__ >= 0 && dev__ < 64) raised__[dev__] = true;                                              \
        }                                                                                                      \
    } while (0)

static inline size_t skb_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline bool skb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// volume limits shared by every entry point (see include/skoots_b200.h)
int skb_check_volume(int64_t X, int64_t Y, int64_t Z, const char* who);

// ---- element conversion -------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float skb_to_float(T v);
template <> __device__ __forceinline__ float skb_to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float skb_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float skb_to_float<float>(float v) { return v; }

template <typename T> __device__ __forceinline__ T skb_from_float(float v);
template <> __device__ __forceinline__ __half skb_from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 skb_from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ float skb_from_float<float>(float v) { return v; }

// ---- streaming loads/stores: bypass L1 for data touched once ---------------------------------
__device__ __forceinline__ uint4 skb_ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void skb_st_stream16(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// ---- sparse CCL workspace layout (skb_ccl.cu writes it, skb_assemble.cu reads it) ---------------
struct SkbCclHeader {
    unsigned n_tile_roots;    // appended by the tile kernel
    unsigned n_global_roots;  // appended by the flatten kernel
    int n_components;         // written by the scan kernel
    int label_base;
    int planar;
    int capacity;
    int dims[3];
    int reserved[7];
};

constexpr int SKB_SCAN_TILE = 8192;
constexpr int SKB_TILE_CURSORS = 256;        // work cursors of the tile kernel, one 128-byte line each
constexpr int SKB_TILE_CURSOR_STRIDE = 32;   // ints

struct SkbCclLayout {
    int X, Y, Z, ZW;      // ZW = 64-bit words per (x,y) row of the bit-packed mask
    int64_t V, n_words, n_chunks;
    int64_t n_scan_tiles;  // chunk histogram is scanned in tiles of SKB_SCAN_TILE entries
    size_t off_bits, off_parent, off_rootbits, off_chunks, off_scan_tiles, off_face_lo, off_face_hi, off_cursors,
        off_tile_roots, off_flat, off_groots, total;
};

static inline SkbCclLayout skb_ccl_layout(int64_t X, int64_t Y, int64_t Z, int64_t capacity) {
    SkbCclLayout L;
    L.X = (int)X; L.Y = (int)Y; L.Z = (int)Z;
    L.ZW = (int)((Z + 63) / 64);
    L.V = X * Y * Z;
    L.n_words = X * Y * (int64_t)L.ZW;
    L.n_chunks = (L.n_words + 63) / 64;
    size_t at = 256;  // header
    L.off_bits = at;       at = skb_align_up(at + (size_t)L.n_words * 8, 256);
    L.off_parent = at;     at = skb_align_up(at + (size_t)L.V * 4, 256);
    L.off_rootbits = at;   at = skb_align_up(at + (size_t)L.n_words * 8, 256);
    L.off_chunks = at;     at = skb_align_up(at + (size_t)(L.n_chunks + 1) * 4, 256);
    L.n_scan_tiles = (L.n_chunks + SKB_SCAN_TILE - 1) / SKB_SCAN_TILE;
    L.off_scan_tiles = at; at = skb_align_up(at + (size_t)(L.n_scan_tiles + 1) * 4, 256);
    // sharded mode: compact copies of every row's first / last word of the slab (the planes the neighbours need)
    L.off_face_lo = at;    at = skb_align_up(at + (size_t)X * Y * 8, 256);
    L.off_face_hi = at;    at = skb_align_up(at + (size_t)X * Y * 8, 256);
    L.off_cursors = at;    at = skb_align_up(at + (size_t)SKB_TILE_CURSORS * SKB_TILE_CURSOR_STRIDE * 4, 256);
    L.off_tile_roots = at; at = skb_align_up(at + (size_t)capacity * 4, 256);
    L.off_flat = at;       at = skb_align_up(at + (size_t)capacity * 4, 256);
    L.off_groots = at;     at = skb_align_up(at + (size_t)capacity * 4, 256);
    L.total = at;
    return L;
}

// ---- sharded pass: per-rank mailbox in peer-visible memory (skb_peer.cu allocates it) -----------------
// Every rank uses the same layout, so a rank computes the address of a slot in a peer's mailbox from
// the peer's base pointer alone.  Two copies ("parities") of every receive buffer: pass k uses copy
// k & 1, which is what makes a peer that runs one pass ahead harmless.
constexpr int SKB_FLAG_STRIDE = 32;  // ints between two flag words (one 128-byte line each)

struct SkbMailboxLayout {
    int world, cap_runs, cap_roots, cap_pairs;
    long long stride;         // ints of one rank's gather payload: [n_roots, n_pairs, roots, pairs]
    long long runs_ints;      // ints of one boundary-run buffer: [count,_,_] + triples
    size_t off_epoch;         // int  pass counter (local)
    size_t off_cnt;           // int[2] run counters of my low / high face (local)
    size_t off_flag_lo;       // int  = k once the lower neighbour's runs of pass k are in recv_lo[k & 1]
    size_t off_flag_hi;       // int  same for the upper neighbour / recv_hi
    size_t off_flag_gather;   // int[world] (SKB_FLAG_STRIDE apart) = k once rank p's payload of pass k is in gathered[k & 1][p]
    size_t off_recv_lo, off_recv_hi;  // 2 x runs_ints ints each
    size_t off_gathered;      // 2 x world x stride ints
    size_t total;
};

static inline SkbMailboxLayout skb_mailbox_layout(int world, int64_t cap_runs, int64_t cap_roots, int64_t cap_pairs) {
    SkbMailboxLayout M;
    M.world = world; M.cap_runs = (int)cap_runs; M.cap_roots = (int)cap_roots; M.cap_pairs = (int)cap_pairs;
    M.stride = 2 + cap_roots + 2 * cap_pairs;
    M.runs_ints = 3 * (cap_runs + 1);
    size_t at = 0;
    M.off_epoch = at;       at += 256;
    M.off_cnt = at;         at += 256;
    M.off_flag_lo = at;     at += 256;
    M.off_flag_hi = at;     at += 256;
    M.off_flag_gather = at; at = skb_align_up(at + (size_t)world * SKB_FLAG_STRIDE * 4, 256);
    M.off_recv_lo = at;     at = skb_align_up(at + 2 * (size_t)M.runs_ints * 4, 256);
    M.off_recv_hi = at;     at = skb_align_up(at + 2 * (size_t)M.runs_ints * 4, 256);
    M.off_gathered = at;    at = skb_align_up(at + 2 * (size_t)world * (size_t)M.stride * 4, 256);
    M.total = at;
    return M;
}

// label of a foreground voxel `t` from the sparse form: parent[t] is either the (negative) label
// code or the index of the voxel's tile root, whose entry is the code.
__device__ __forceinline__ int skb_sparse_label(const int* __restrict__ parent, int t) {
    int p = __ldg(parent + t);
    if (p >= 0) p = __ldg(parent + p);
    return -p;
}
