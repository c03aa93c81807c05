// Connected-component labelling of the thresholded skeleton mask on B200 (sm_100a).
//
// Replaces the arithmetic of skoots/lib/flood_fill.py:13-140 (scipy.ndimage.label per crop +
// seam merging).  Design:
//
//   K1a ccl_pack_kernel     streams the mask once (16-byte loads, Z is the contiguous axis) into the
//                           bit-packed mask: one 64-bit word per 64 z of an (x,y) row.
//   K1b ccl_tile_kernel     one WARP per 8x8x64 tile of the bit mask: empty tiles cost two loads and
//                           a vote; otherwise a warp-synchronous shared-memory union-find over
//                           z-RUNS (a run's start is found with clz on the row word, so
//                           z-connectivity costs nothing; only y/x neighbour rows need unions).
//                           Every foreground voxel gets parent[v] = its tile root; tile roots
//                           are appended to a list.
//   K2 ccl_boundary_kernel  one thread per 64-bit word of the bit-packed mask: unions across
//                           tile faces with atomicMin on the global parent array.
//   K3 ccl_flatten_kernel   pointer-jumps every tile root to its global root, marks global
//                           roots in a bitmap and counts them per 4096-voxel chunk.
//   K4 ccl_scan_*_kernel    two-level exclusive scan of the chunk counts (-> raster-order rank bases).
//   K5 ccl_rank_kernel      rank of each global root = scipy.ndimage.label's numbering.
//   K6 ccl_publish_kernel   copies the label code into every non-global tile root.
//
// After K6:  parent[v] < 0  -> label = -parent[v]      (v is a tile root)
//            parent[v] >= 0 -> label = -parent[parent[v]]
// Only foreground entries of `parent` are ever written or read, so the dense 4-byte array is
// never initialised and never streamed: HBM traffic is 1 B/voxel (mask) + 1/8 B (bit mask)
// + O(foreground).
#include <stdarg.h>
#include <stdio.h>

#include <stdlib.h>

#include "skb_ccl.cuh"

// ------------------------------------------------------------------------------------------
// error plumbing shared by every translation unit
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void skb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* skb_last_error(void) { return g_err; }
extern "C" int skb_version(void) { return SKB_VERSION; }

int skb_check_volume(int64_t X, int64_t Y, int64_t Z, const char* who) {
    if (X <= 0 || Y <= 0 || Z <= 0) {
        skb_set_error("%s: dims must be positive (got %lld,%lld,%lld)", who, (long long)X, (long long)Y, (long long)Z);
        return SKB_E_ARG;
    }
    if (X >= (1 << 24) || Y >= (1 << 24) || Z >= (1 << 24) || X * Y * Z > 2147483648LL) {
        skb_set_error("%s: volume %lldx%lldx%lld exceeds 2^31 voxels / 2^24 per axis", who, (long long)X,
                      (long long)Y, (long long)Z);
        return SKB_E_RANGE;
    }
    return SKB_OK;
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
// start of the z-run containing bit p of row word w (bit p must be set)
__device__ __forceinline__ int run_start(ull w, int p) {
    ull below = ~w & ((1ull << p) - 1ull);
    return below ? 64 - __clzll((long long)below) : 0;
}

__device__ __forceinline__ int sfind(volatile int* lab, int a) {
    int p = lab[a];
    while (p != a) {
        a = p;
        p = lab[a];
    }
    return a;
}

__device__ __forceinline__ void sunion(int* lab, int a, int b) {
    for (;;) {
        a = sfind(lab, a);
        b = sfind(lab, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(&lab[a], b);
        if (old == a) return;
        a = old;
    }
}

// ---- mask elements -> foreground bits ---------------------------------------------------------
// 4 bytes -> 4 bits (byte != 0).  High bit of every non-zero byte, then a multiply gathers the four
// flags into one nibble (the partial products land on distinct bit positions, so no carries).
__device__ __forceinline__ unsigned nz4(unsigned w) {
    unsigned t = (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
    return (((t >> 7) * 0x01020408u) >> 24) & 0xFu;
}
// 2 int16 -> 2 bits (element > 0): non-zero and sign bit clear
__device__ __forceinline__ unsigned gt2(unsigned w) {
    unsigned t = (((w & 0x7fff7fffu) + 0x7fff7fffu) | w) & ~w & 0x80008000u;
    return ((t >> 15) | (t >> 30)) & 3u;
}

// foreground bits of one lane's segment of SEG (= TZ/4) consecutive mask elements; n of them exist.
// Vector path: one 4/8/16/32-byte load, an all-zero early-out (the common case in a sparse mask),
// then 4 (u8) or 2 (i16) elements per 32-bit word through nz4 / gt2.
template <typename MaskT, int SEG>
__device__ __forceinline__ unsigned seg_bits(const MaskT* p, int n, bool vec_ok) {
    constexpr int BYTES = SEG * (int)sizeof(MaskT);
    constexpr int NWORD = BYTES / 4;
    constexpr int PER = 4 / (int)sizeof(MaskT);  // elements per 32-bit word
    unsigned bits = 0;
    if (vec_ok && n == SEG) {
        unsigned w[NWORD];
        if (BYTES == 4) {
            w[0] = __ldg(reinterpret_cast<const unsigned*>(p));
        } else if (BYTES == 8) {
            uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
            w[0] = q.x; w[1] = q.y;
        } else {
#pragma unroll
            for (int k = 0; k < BYTES / 16; ++k) {
                uint4 q = skb_ld_stream16(reinterpret_cast<const uint4*>(p) + k);
                w[4 * k] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w;
            }
        }
        unsigned any = 0;
#pragma unroll
        for (int k = 0; k < NWORD; ++k) any |= w[k];
        if (any) {
#pragma unroll
            for (int k = 0; k < NWORD; ++k) bits |= (sizeof(MaskT) == 1 ? nz4(w[k]) : gt2(w[k])) << (PER * k);
        }
    } else {
        for (int i = 0; i < n; ++i) bits |= (unsigned)(p[i] > 0) << i;
    }
    return bits;
}

__global__ void ccl_init_kernel(CclView v, SkbCclHeader h, int keep_status) {  // <<<1, SKB_TILE_CURSORS>>>
    v.cursors[threadIdx.x * SKB_TILE_CURSOR_STRIDE] = 0u;
    if (threadIdx.x == 0) {
        *v.hdr = h;
        if (!keep_status) *v.status = 0u;
        if (v.ncomp_out) *v.ncomp_out = 0;
    }
}

// ------------------------------------------------------------------------------------------
// K1: tile-local run-based union-find
// ------------------------------------------------------------------------------------------
// K1a  mask -> bit-packed mask: a pure stream.  Every lane reads 16 consecutive mask elements with
// one (u8) or two (i16) 16-byte loads, turns them into 16 bits (all-zero early-out), four lanes
// combine their bits into one 64-bit word with two shuffles and lane 0 of the group stores it:
// a warp reads 512 contiguous bytes and writes 64 contiguous bytes.
constexpr int PACK_SEGS = 4;

template <typename MaskT>
__global__ void __launch_bounds__(256) ccl_pack_kernel(const MaskT* __restrict__ mask, CclView v, unsigned n_seg,
                                                      int nk_shift) {
    // PACK_SEGS 16-element segments per thread (256 segments apart: every load coalesced, all in flight together)
    const unsigned t0 = blockIdx.x * (256u * PACK_SEGS) + threadIdx.x;
    unsigned b[PACK_SEGS];
    uint4 raw[PACK_SEGS][sizeof(MaskT)];
#pragma unroll
    for (int u = 0; u < PACK_SEGS; ++u) {
        const unsigned t = t0 + 256u * u;
#pragma unroll
        for (int h = 0; h < (int)sizeof(MaskT); ++h)
            raw[u][h] = t < n_seg ? skb_ld_stream16(reinterpret_cast<const uint4*>(mask + (size_t)t * 16) + h) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < PACK_SEGS; ++u) {
        unsigned bits = 0;
#pragma unroll
        for (int h = 0; h < (int)sizeof(MaskT); ++h) {
            const unsigned w[4] = {raw[u][h].x, raw[u][h].y, raw[u][h].z, raw[u][h].w};
            if (w[0] | w[1] | w[2] | w[3]) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    bits |= (sizeof(MaskT) == 1 ? nz4(w[k]) : gt2(w[k])) << ((sizeof(MaskT) == 1 ? 4 : 2) * (4 * h + k));
            }
        }
        b[u] = bits;
    }
#pragma unroll
    for (int u = 0; u < PACK_SEGS; ++u) {
        const unsigned t = t0 + 256u * u;
        ull w = (ull)b[u] << (16 * (t & 3u));
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        if ((t & 3u) == 0u && t < n_seg) {
            const unsigned wl = t >> 2;  // word of the slab, flat order = (row, k local) = its place in `bits`
            if (v.face_lo) {
                const unsigned rowi = nk_shift >= 0 ? (wl >> nk_shift) : wl / (unsigned)v.nk;
                const unsigned kl = wl - rowi * (unsigned)v.nk;
                if (kl == 0u) v.face_lo[rowi] = w;
                if (kl == (unsigned)v.nk - 1u) v.face_hi[rowi] = w;
            }
            v.bits[wl] = w;
        }
    }
}

// any Z / unaligned masks: one thread per word, element-wise
template <typename MaskT>
__global__ void __launch_bounds__(256) ccl_pack_generic_kernel(const MaskT* __restrict__ mask, CclView v) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned rows = (unsigned)v.X * (unsigned)v.Y;
    const unsigned rowi = t / (unsigned)v.nk, kl = t - rowi * (unsigned)v.nk;
    if (rowi >= rows) return;
    const int zl0 = (int)kl * 64, n = min(64, v.Zl - zl0);
    const MaskT* p = mask + (size_t)rowi * v.Zl + zl0;
    ull w = 0;
    for (int i = 0; i < n; ++i) w |= (ull)(p[i] > 0) << i;
    v.bits[t] = w;
    if (v.face_lo) {
        if (kl == 0u) v.face_lo[rowi] = w;
        if (kl == (unsigned)v.nk - 1u) v.face_hi[rowi] = w;
    }
}

// 16-bit labels in shared memory: atomic min through a 32-bit CAS; returns the previous value
__device__ __forceinline__ int amin16(unsigned short* lab, int idx, int val) {
    unsigned* w = reinterpret_cast<unsigned*>(lab) + (idx >> 1);
    const int sh = (idx & 1) * 16;
    unsigned old = *reinterpret_cast<volatile unsigned*>(w);
    for (;;) {
        const int cur = (int)((old >> sh) & 0xffffu);
        if (cur <= val) return cur;
        const unsigned nw = (old & ~(0xffffu << sh)) | ((unsigned)val << sh);
        const unsigned prev = atomicCAS(w, old, nw);
        if (prev == old) return cur;
        old = prev;
    }
}

__device__ __forceinline__ int sfind16(const volatile unsigned short* lab, int a) {
    int p = lab[a];
    while (p != a) {
        a = p;
        p = lab[a];
    }
    return a;
}

__device__ __forceinline__ void sunion16(unsigned short* lab, int a, int b) {
    for (;;) {
        a = sfind16(lab, a);
        b = sfind16(lab, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = amin16(lab, a, b);
        if (old == a) return;
        a = old;
    }
}

// K1b  one WARP per 8 x 8 x 64 tile, working on the bit-packed mask (8x less data than the mask and no
// CTA-wide barriers): each lane owns two of the tile's 64 row words.  An empty tile costs two 8-byte
// loads and a vote.  Otherwise a warp-synchronous union-find over z-runs: a run's start is found
// with clz on the row word, so z-connectivity costs nothing; a run start can only sit at every
// other bit, so 32 slots per row suffice: slot = row*32 + (p>>1), 16-bit labels.
__device__ __forceinline__ void tile_row_unions(unsigned short* slab, const ull* srow, int row, ull w, bool connect_x) {
    if (!w) return;
    if (row & 7) {  // y-1 inside the tile
        ull wn = srow[row - 1], a = w & wn;
        for (ull s = a & ~(a << 1); s; s &= s - 1) {
            int p = __ffsll((long long)s) - 1;
            sunion16(slab, row * 32 + (run_start(w, p) >> 1), (row - 1) * 32 + (run_start(wn, p) >> 1));
        }
    }
    if ((row >> 3) && connect_x) {  // x-1 inside the tile
        ull wn = srow[row - 8], a = w & wn;
        for (ull s = a & ~(a << 1); s; s &= s - 1) {
            int p = __ffsll((long long)s) - 1;
            sunion16(slab, row * 32 + (run_start(w, p) >> 1), (row - 8) * 32 + (run_start(wn, p) >> 1));
        }
    }
}

__device__ __forceinline__ ull tile_row_resolve(const CclView& v, unsigned short* slab, const ull* srow, int row, ull w,
                                                int x0, int y0, int z0, int& gbase_out) {
    gbase_out = 0;
    if (!w) return 0ull;
    const unsigned rowi = (unsigned)(x0 + (row >> 3)) * (unsigned)v.Y + (unsigned)(y0 + (row & 7));
    const int gbase = (int)(rowi * (unsigned)v.Z) + z0;  // voxel index of bit 0 of this row word
    ull rootmask = 0;                                    // bit p set when the run starting at p is a tile root
    for (ull s = w & ~(w << 1); s; s &= s - 1) {
        const int p = __ffsll((long long)s) - 1;
        const int l = row * 32 + (p >> 1);
        const int r = sfind16(slab, l);
        int groot;
        if (r == l) {
            rootmask |= 1ull << p;
            groot = gbase + p;
        } else {
            const int rrow = r >> 5, rh = r & 31;
            const ull rs = srow[rrow] & ~(srow[rrow] << 1);
            const int rp = 2 * rh + (int)((rs >> (2 * rh + 1)) & 1ull);
            groot = (int)(((unsigned)(x0 + (rrow >> 3)) * (unsigned)v.Y + (unsigned)(y0 + (rrow & 7))) * (unsigned)v.Z) + z0 + rp;
        }
        const ull tt = ~(w >> p);
        const int len = tt ? __ffsll((long long)tt) - 1 : 64 - p;
        for (int j = 0; j < len; ++j) v.parent[gbase + p + j] = groot;
    }
    gbase_out = gbase;
    return rootmask;
}

// Tile roots are collected in a per-warp shared-memory buffer and appended to the global list with ONE
// atomicAdd per flush: a counter that every non-empty tile bumps individually serialises in L2
// (hundreds of thousands of same-address atomics cost more than streaming the whole mask).
constexpr int CCL_ROOT_BUF = 128;

__device__ __forceinline__ void flush_roots(const CclView& v, int* sbuf, int& buf_n, int lane) {
    if (buf_n == 0) return;
    int base = 0;
    if (lane == 0) base = (int)atomicAdd(&v.hdr->n_tile_roots, (unsigned)buf_n);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int i = lane; i < buf_n; i += 32) {
        if (base + i < v.capacity) v.tile_roots[base + i] = sbuf[i];
        else atomicOr(v.status, SKB_STATUS_ROOT_OVERFLOW);
    }
    __syncwarp();
    buf_n = 0;
}

// warp-collective: every lane passes the root masks of its two rows
__device__ __forceinline__ void append_roots(const CclView& v, int* sbuf, int& buf_n, int lane, ull m0, int g0, ull m1, int g1) {
    const int cnt = __popcll(m0) + __popcll(m1);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    if (total > CCL_ROOT_BUF) {  // very dense tile: straight to the list
        flush_roots(v, sbuf, buf_n, lane);
        int base = 0;
        if (lane == 0) base = (int)atomicAdd(&v.hdr->n_tile_roots, (unsigned)total);
        int at = __shfl_sync(0xffffffffu, base, 0) + incl - cnt;
        for (ull m = m0; m; m &= m - 1, ++at) {
            if (at < v.capacity) v.tile_roots[at] = g0 + __ffsll((long long)m) - 1;
            else atomicOr(v.status, SKB_STATUS_ROOT_OVERFLOW);
        }
        for (ull m = m1; m; m &= m - 1, ++at) {
            if (at < v.capacity) v.tile_roots[at] = g1 + __ffsll((long long)m) - 1;
            else atomicOr(v.status, SKB_STATUS_ROOT_OVERFLOW);
        }
        return;
    }
    if (buf_n + total > CCL_ROOT_BUF) flush_roots(v, sbuf, buf_n, lane);
    int at = buf_n + incl - cnt;
    for (ull m = m0; m; m &= m - 1) sbuf[at++] = g0 + __ffsll((long long)m) - 1;
    for (ull m = m1; m; m &= m - 1) sbuf[at++] = g1 + __ffsll((long long)m) - 1;
    buf_n += total;
    __syncwarp();
}

// ---- sparse tiles: one z-run per lane ---------------------------------------------------------------
// A skeleton tile holds ~15 runs in 64 rows.  With a row per lane (the dense path below) most lanes idle
// while the others walk their rows' runs one after the other through divergent loops: ncu counted ~2050
// warp instructions per non-empty tile at 10 active threads.  Here the tile's runs are first numbered in
// raster order (row, then z) and listed in shared memory, then each lane takes ONE run: it unions with the
// runs it touches in rows y-1 and x-1 (the neighbour's run number = the row's first number + the rank of
// the run among the row's run starts, a popcount) on 32-bit labels with native shared-memory atomicMin,
// and finally writes its run's voxels and reports itself if it is a root.  Lowest number = first run in
// raster order, so a tile root is still the first voxel of its tile-component.
constexpr int CCL_MAX_RUNS = 96;  // more runs than this in a tile -> dense path

__device__ __forceinline__ int lfind(const volatile int* lab, int a) {
    int p = lab[a];
    while (p != a) {
        a = p;
        p = lab[a];
    }
    return a;
}

__device__ __forceinline__ void lunion(int* lab, int a, int b) {
    for (;;) {
        a = lfind(lab, a);
        b = lfind(lab, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(&lab[a], b);
        if (old == a) return;
        a = old;
    }
}

struct SparseTileScratch {  // aliases the dense path's label array (a tile takes one path or the other)
    unsigned short rowbase[64];         // number of the row's first run
    unsigned short run[CCL_MAX_RUNS];   // row << 6 | start bit
    int lab[CCL_MAX_RUNS];
};

__device__ __forceinline__ void link_row(const ull* srow, const SparseTileScratch* sc, int* lab, int i, ull mask, int nrow) {
    const ull wn = srow[nrow], a = mask & wn;
    if (!a) return;
    const ull stn = wn & ~(wn << 1);
    for (ull s = a & ~(a << 1); s; s &= s - 1) {
        const int b = __ffsll((long long)s) - 1;  // first voxel of an overlap: which run of the neighbour row holds it?
        const int j = (int)sc->rowbase[nrow] + __popcll(stn & ((2ull << b) - 1ull)) - 1;
        lunion(lab, i, j);
    }
}

// returns false (and does nothing) if the tile has more than CCL_MAX_RUNS runs
__device__ __forceinline__ bool tile_sparse(const CclView& v, ull* srow, SparseTileScratch* sc, int* sbuf, int& buf_n,
                                            int lane, ull w0, ull w1, int x0, int y0, int z0) {
    const ull st0 = w0 & ~(w0 << 1), st1 = w1 & ~(w1 << 1);
    const int c0 = __popcll(st0), c1 = __popcll(st1);
    int incl = c0 | (c1 << 16);  // both row sets in one scan (a set holds at most 32 x 32 runs)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    const int tot0 = tot & 0xffff, R = tot0 + (tot >> 16);
    if (R > CCL_MAX_RUNS) return false;
    const int base0 = (incl & 0xffff) - c0, base1 = tot0 + (incl >> 16) - c1;
    srow[lane] = w0;
    srow[lane + 32] = w1;
    sc->rowbase[lane] = (unsigned short)base0;
    sc->rowbase[lane + 32] = (unsigned short)base1;
    int at = base0;
    for (ull s = st0; s; s &= s - 1, ++at) {
        sc->run[at] = (unsigned short)((lane << 6) | (__ffsll((long long)s) - 1));
        sc->lab[at] = at;
    }
    at = base1;
    for (ull s = st1; s; s &= s - 1, ++at) {
        sc->run[at] = (unsigned short)(((lane + 32) << 6) | (__ffsll((long long)s) - 1));
        sc->lab[at] = at;
    }
    __syncwarp();
    // unions: one run per lane
    for (int i = lane; i < R; i += 32) {
        const int e = sc->run[i], row = e >> 6, p = e & 63;
        const ull w = srow[row], tt = ~(w >> p);
        const int len = tt ? __ffsll((long long)tt) - 1 : 64 - p;
        const ull mask = (len >= 64 ? ~0ull : ((1ull << len) - 1ull)) << p;
        if (row & 7) link_row(srow, sc, sc->lab, i, mask, row - 1);
        if ((row >> 3) && v.connect_x) link_row(srow, sc, sc->lab, i, mask, row - 8);
    }
    __syncwarp();
    // resolve: every run's voxels point at the tile root's first voxel; roots are listed
    for (int base = 0; base < R; base += 32) {
        const int i = base + lane;
        bool is_root = false;
        int groot = 0;
        if (i < R) {
            const int r = lfind(sc->lab, i);
            const int er = sc->run[r], rrow = er >> 6;
            groot = (int)(((unsigned)(x0 + (rrow >> 3)) * (unsigned)v.Y + (unsigned)(y0 + (rrow & 7))) * (unsigned)v.Z) + z0 + (er & 63);
            is_root = r == i;
            const int e = sc->run[i], row = e >> 6, p = e & 63;
            const ull tt = ~(srow[row] >> p);
            const int len = tt ? __ffsll((long long)tt) - 1 : 64 - p;
            int* dst = v.parent + (int)(((unsigned)(x0 + (row >> 3)) * (unsigned)v.Y + (unsigned)(y0 + (row & 7))) * (unsigned)v.Z) + z0 + p;
            for (int j = 0; j < len; ++j) dst[j] = groot;
        }
        const unsigned m = __ballot_sync(0xffffffffu, is_root);
        const int n = __popc(m);
        if (n) {
            if (buf_n + n > CCL_ROOT_BUF) flush_roots(v, sbuf, buf_n, lane);
            if (is_root) sbuf[buf_n + __popc(m & ((1u << lane) - 1u))] = groot;
            buf_n += n;
            __syncwarp();
        }
    }
    __syncwarp();  // shared memory is reused by this warp's next tile
    return true;
}

constexpr int CCL_TILE_WARPS = 8;
constexpr int CCL_TILE_BATCH = 4;  // consecutive tiles a warp claims at a time

// Warps are PERSISTENT and independent (with one tile per warp and 8 warps per CTA the CTA's shared
// memory stays pinned until its slowest warp — the one tile in four that is not empty — has finished).
// Work is handed out DYNAMICALLY, a batch of CCL_TILE_BATCH consecutive tiles per atomicAdd: a non-empty
// tile costs ~50x an empty one and a slab of an 8-way split has only ~11 tiles per warp, so a static
// assignment finishes with the unluckiest warp (measured: 1.7x the per-voxel time of the whole
// volume).  ONE cursor would not do — same-address atomics retire at ~20 ns each on B200 (measured:
// 16 K claims on one word turned a 110 us kernel into 367 us) — so the batch list is cut into
// SKB_TILE_CURSORS interleaved ranges with a cursor each (own 128-byte line); a warp serves the range
// (its global id mod SKB_TILE_CURSORS) and, once that is exhausted, steals from the next range that
// still has work.  It claims its next batch a whole batch ahead of its use.
__global__ void __launch_bounds__(32 * CCL_TILE_WARPS, 5) ccl_tile_kernel(CclView v, unsigned n_tiles, unsigned n_yk,
                                                                       int nk_shift, int nyk_shift, int dynamic) {
    __shared__ ull srow_all[CCL_TILE_WARPS][64];
    __shared__ __align__(16) unsigned short slab_all[CCL_TILE_WARPS][64 * 32];
    __shared__ int rootbuf_all[CCL_TILE_WARPS][CCL_ROOT_BUF];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* sbuf = rootbuf_all[warp];
    int buf_n = 0;
    ull* srow = srow_all[warp];
    unsigned short* slab = slab_all[warp];
    const int r0 = lane, r1 = lane + 32;
    // dynamic: batches of CCL_TILE_BATCH consecutive tiles claimed from the cursors.  static (few tiles per warp,
    // where the claims cost more than the imbalance — measured 141 vs 110 us on a 1/8 slab): single tiles, strided.
    const unsigned tile_batch = dynamic ? CCL_TILE_BATCH : 1;
    const unsigned n_batches = (n_tiles + tile_batch - 1) / tile_batch;
    const unsigned n_warps = gridDim.x * CCL_TILE_WARPS;
    unsigned static_next = blockIdx.x * CCL_TILE_WARPS + warp;

    struct Tile { int x0, y0, k; };
    auto decode = [&](unsigned t) -> Tile {  // tile list order: k fastest, then y tile, then x tile
        const unsigned xt = nyk_shift >= 0 ? (t >> nyk_shift) : t / n_yk, yk = t - xt * n_yk;
        const unsigned yt = nk_shift >= 0 ? (yk >> nk_shift) : yk / (unsigned)v.nk;
        return Tile{(int)xt * 8, (int)yt * 8, v.k0 + (int)(yk - yt * (unsigned)v.nk)};
    };
    auto load = [&](const Tile& T, ull& w0, ull& w1) {
        const int xa = T.x0 + (r0 >> 3), xb = T.x0 + (r1 >> 3), y = T.y0 + (lane & 7);
        w0 = 0; w1 = 0;
        if (y < v.Y) {
            if (xa < v.X) w0 = v.bits[((size_t)xa * v.Y + y) * v.nk + (T.k - v.k0)];
            if (xb < v.X) w1 = v.bits[((size_t)xb * v.Y + y) * v.nk + (T.k - v.k0)];
        }
    };
    // range r owns the batches r, r + n_ranges, r + 2 n_ranges, ... (strided, so that every range samples the
    // whole volume and the ranges carry about the same work); cursor r counts how many of them are claimed
    const unsigned n_ranges = min((unsigned)SKB_TILE_CURSORS, gridDim.x * CCL_TILE_WARPS);  // every range has a warp
    unsigned my = (blockIdx.x * CCL_TILE_WARPS + warp) % n_ranges;
    auto claim = [&](unsigned r) -> unsigned {  // lane 0's value is the one that counts
        return lane == 0 ? r + n_ranges * atomicAdd(v.cursors + r * SKB_TILE_CURSOR_STRIDE, 1u) : 0u;
    };
    unsigned pending = dynamic ? claim(my) : 0u, scanned = 0;
    // the batch claimed last, or, when my range is exhausted, one stolen from the next range that still has work
    // (exhausted cursors are skipped with plain loads: only a cursor that looks alive costs an atomic)
    auto next_batch = [&](unsigned& batch) -> bool {
        if (!dynamic) {
            batch = static_next;
            static_next += n_warps;
            return batch < n_batches;
        }
        for (;;) {
            batch = __shfl_sync(0xffffffffu, pending, 0);
            if (batch < n_batches) {
                pending = claim(my);  // a whole batch ahead of its use
                return true;
            }
            unsigned found = n_ranges;
            while (scanned < n_ranges && found == n_ranges) {
                const unsigned off = scanned + 1u + lane;
                unsigned r = my + off;
                r -= r >= n_ranges ? n_ranges : 0u;
                bool alive = false;
                if (off < n_ranges) {
                    const unsigned c = *reinterpret_cast<volatile unsigned*>(v.cursors + r * SKB_TILE_CURSOR_STRIDE);
                    alive = (unsigned long long)r + (unsigned long long)n_ranges * c < n_batches;
                }
                const unsigned m = __ballot_sync(0xffffffffu, alive);
                if (m) {
                    const int first = __ffs((int)m) - 1;
                    found = __shfl_sync(0xffffffffu, r, first);
                    scanned += (unsigned)first;  // the ranges before it are exhausted for good
                } else {
                    scanned += 32u;
                }
            }
            if (found == n_ranges) return false;
            my = found;
            scanned = 0;
            pending = claim(my);
        }
    };

    // Measured and NOT adopted (round 2): a tile-occupancy bitmap written by the pack kernel so that this kernel would not load
    // the words of empty tiles.  The tile kernel went from 244 to 202 us on the headline volume, but marking the tiles cost the
    // pack kernel — a pure stream at the HBM roofline — far more: 340 -> ~440 us with one atomicOr per non-empty word (8 M
    // atomics on 16 K words), 793 us with a read-before-atomic (profiles/r02_launches_ccl_tile_bitmap_experiment.txt).
    // One tile at a time, the next tile's two row words in flight while the current one is labelled (a
    // single copy of the labelling code: unrolling it over a batch quadrupled the kernel's time).
    unsigned batch;
    if (!next_batch(batch)) return;
    unsigned t = batch * tile_batch, t_end = min(t + tile_batch, n_tiles);
    Tile cur = decode(t);
    ull w0, w1;
    load(cur, w0, w1);
    for (;;) {
        unsigned tn = t + 1;
        bool more = true;
        if (tn >= t_end) {
            more = next_batch(batch);
            if (more) {
                tn = batch * tile_batch;
                t_end = min(tn + tile_batch, n_tiles);
            }
        }
        Tile nxt = cur;
        ull n0 = 0, n1 = 0;
        if (more) {
            nxt = decode(tn);
            load(nxt, n0, n1);
        }
        if (__any_sync(0xffffffffu, (w0 | w1) != 0ull) &&
            !tile_sparse(v, srow, reinterpret_cast<SparseTileScratch*>(slab), sbuf, buf_n, lane, w0, w1, cur.x0, cur.y0, 64 * cur.k)) {
            // dense tile: a row per lane, 16-bit labels indexed by (row, run start / 2)
            const int z0 = 64 * cur.k;
            srow[r0] = w0;
            srow[r1] = w1;
            for (ull s = w0 & ~(w0 << 1); s; s &= s - 1) {
                const int h = (__ffsll((long long)s) - 1) >> 1;
                slab[r0 * 32 + h] = (unsigned short)(r0 * 32 + h);
            }
            for (ull s = w1 & ~(w1 << 1); s; s &= s - 1) {
                const int h = (__ffsll((long long)s) - 1) >> 1;
                slab[r1 * 32 + h] = (unsigned short)(r1 * 32 + h);
            }
            __syncwarp();
            tile_row_unions(slab, srow, r0, w0, v.connect_x != 0);
            tile_row_unions(slab, srow, r1, w1, v.connect_x != 0);
            __syncwarp();
            int g0, g1;
            const ull m0 = tile_row_resolve(v, slab, srow, r0, w0, cur.x0, cur.y0, z0, g0);
            const ull m1 = tile_row_resolve(v, slab, srow, r1, w1, cur.x0, cur.y0, z0, g1);
            __syncwarp();  // shared memory is reused by this warp's next tile
            append_roots(v, sbuf, buf_n, lane, m0, g0, m1, g1);
        }
        if (!more) break;
        t = tn; cur = nxt; w0 = n0; w1 = n1;
    }
    flush_roots(v, sbuf, buf_n, lane);
}

// ------------------------------------------------------------------------------------------
// K2: unions across tile faces, driven by the bit-packed mask
// ------------------------------------------------------------------------------------------
// What one word of the bit mask has to union across tile faces: the run starts (per face) whose voxel p
// is foreground on both sides.  Only words that contain foreground get here.  n_words <= 2^31: 32-bit
// index math, shifts when the row length / Y are powers of two (the usual case).
// widx = place of the word in `bits` = row * nk + (word of the slab's row).
struct FaceWork {
    ull sy, sx;   // bit p set: union voxel gbase+p with its y-1 / x-1 neighbour
    int z;        // 1: union voxel gbase with gbase-1 (run continues from the previous word of the row)
    int gbase;
    __device__ __forceinline__ int count() const { return z + __popcll(sy) + __popcll(sx); }
};

__device__ __forceinline__ FaceWork boundary_word(const CclView& v, unsigned widx, ull w, bool have_prev, ull prev, int TX, int TY) {
    FaceWork f = {0ull, 0ull, 0, 0};
    if (!w) return f;
    const unsigned uzw = (unsigned)v.nk, uyy = (unsigned)v.Y;
    unsigned rowi, k, x, y;
    if (v.nk_shift >= 0) { rowi = widx >> v.nk_shift; k = widx & (uzw - 1u); }
    else { rowi = widx / uzw; k = widx - rowi * uzw; }
    if (v.y_shift >= 0) { x = rowi >> v.y_shift; y = rowi & (uyy - 1u); }
    else { x = rowi / uyy; y = rowi - x * uyy; }
    const bool face_z = k > 0u && (w & 1ull);
    const bool face_y = y > 0 && (y & (unsigned)(TY - 1)) == 0;  // tile sides are powers of two (8)
    const bool face_x = v.connect_x && x > 0 && (x & (unsigned)(TX - 1)) == 0;
    if (!(face_z || face_y || face_x)) return f;
    f.gbase = (int)(rowi * (unsigned)v.Z + 64u * (k + (unsigned)v.k0));
    // the (up to three) neighbour words are independent loads: issue them together
    ull wz = prev, wy = 0ull, wx = 0ull;
    if (face_z && !have_prev) wz = v.bits[widx - 1];
    if (face_y) wy = v.bits[widx - uzw];
    if (face_x) wx = v.bits[widx - uyy * uzw];
    if (face_z) f.z = (int)(wz >> 63);
    const ull ay = w & wy, ax = w & wx;
    f.sy = ay & ~(ay << 1);
    f.sx = ax & ~(ax << 1);
    return f;
}

template <typename F>
__device__ __forceinline__ void for_each_union(const CclView& v, const FaceWork& f, F&& emit) {
    if (f.z) emit(f.gbase, f.gbase - 1);
    for (ull s = f.sy; s; s &= s - 1) {
        const int p = __ffsll((long long)s) - 1;
        emit(f.gbase + p, f.gbase + p - v.Z);
    }
    const int plane = v.Y * v.Z;
    for (ull s = f.sx; s; s &= s - 1) {
        const int p = __ffsll((long long)s) - 1;
        emit(f.gbase + p, f.gbase + p - plane);
    }
}

// pair_mode: rows hold an even number of words, or one word each (then the two words are two rows and
// there is no z face at all) -> every thread streams two adjacent words with one 16-byte load.
// Otherwise one word per thread.
//
// The unions themselves are pointer chases through `parent` in global memory: ~5 dependent loads of
// ~0.7 us each.  Done where they are found — one lane at a time inside divergent branches — they made
// this kernel latency-bound at 1 active thread per warp (ncu: 69 % long-scoreboard stalls, 52 us for a
// 33 MB bit mask).  So a warp first only COLLECTS its unions into a shared-memory queue and then
// works through the queue one union per lane: 32 chases in flight per warp.
constexpr int BND_WARPS = 8;
constexpr int BND_QUEUE = 96;  // unions a warp can queue; a denser warp falls back to in-place unions

__global__ void __launch_bounds__(32 * BND_WARPS) ccl_boundary_kernel(CclView v, int TX, int TY, int pair_mode) {
    __shared__ int2 s_queue[BND_WARPS][BND_QUEUE];
    const unsigned tix = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int2* queue = s_queue[threadIdx.x >> 5];
    ull w0 = 0ull, w1 = 0ull;
    unsigned widx = 0u;
    if (pair_mode) {
        widx = 2u * tix;
        if (widx < (unsigned)v.n_words) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(v.bits + widx));
            w0 = ((ull)q.y << 32) | q.x;
            w1 = ((ull)q.w << 32) | q.z;
        }
    } else {
        widx = tix;
        if (tix < (unsigned)v.n_words) w0 = v.bits[tix];
    }
    if (!__any_sync(0xffffffffu, (w0 | w1) != 0ull)) return;  // ~90 % of the warps
    const FaceWork f0 = boundary_word(v, widx, w0, false, 0ull, TX, TY);
    const FaceWork f1 = boundary_word(v, widx + 1u, w1, true, w0, TX, TY);
    const int cnt = f0.count() + f1.count();
    if (!__any_sync(0xffffffffu, cnt != 0)) return;  // half of the warps that hold foreground have none on a tile face
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total > BND_QUEUE) {  // dense mask: union in place
        auto direct = [&](int a, int b) { gunion(v.parent, a, b); };
        for_each_union(v, f0, direct);
        for_each_union(v, f1, direct);
        return;
    }
    int at = incl - cnt;
    auto push = [&](int a, int b) { queue[at++] = make_int2(a, b); };
    for_each_union(v, f0, push);
    for_each_union(v, f1, push);
    __syncwarp();
    for (int i = lane; i < total; i += 32) {
        const int2 e = queue[i];
        gunion(v.parent, e.x, e.y);
    }
}

// ------------------------------------------------------------------------------------------
// K3: pointer jumping of tile roots; global roots -> bitmap + chunk histogram + list
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ccl_flatten_kernel(CclView v) {
    // The list of global roots is appended to through ONE counter.  Round 1 took one atomicAdd per warp that held a root:
    // fine for the ~16 K roots of a skeleton volume, but a 64-slice 2-D stack has 1.3 M roots = ~41 K same-address atomics
    // at ~20 ns each (the kernel took 0.36 ms).  Now a CTA gathers its warps' counts in shared memory and takes one global
    // atomic per CTA and iteration.
    __shared__ unsigned s_count, s_base;
    const unsigned n = min(v.hdr->n_tile_roots, (unsigned)v.capacity);
    const unsigned stride = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (unsigned i0 = blockIdx.x * blockDim.x; i0 < n; i0 += stride) {  // uniform per CTA: barriers inside
        const unsigned i = i0 + threadIdx.x;
        bool is_root = false;
        int r = 0;
        if (i < n) {
            r = v.tile_roots[i];
            const int g = gfind(v.parent, r);
            v.flat[i] = g;
            if (g == r) {
                is_root = true;
                int bit;
                long long wi = word_of_voxel(v, r, &bit);
                atomicOr(&v.rootbits[wi], 1ull << bit);
                atomicAdd(&v.chunks[wi >> 6], 1);
            }
        }
        if (threadIdx.x == 0) s_count = 0u;
        __syncthreads();
        const unsigned m = __ballot_sync(0xffffffffu, is_root);
        unsigned wbase = 0u;
        if (lane == 0 && m) wbase = atomicAdd(&s_count, (unsigned)__popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        __syncthreads();
        if (threadIdx.x == 0 && s_count) s_base = atomicAdd(&v.hdr->n_global_roots, s_count);
        __syncthreads();
        if (is_root) v.groots[s_base + wbase + __popc(m & ((1u << lane) - 1u))] = r;
        __syncthreads();  // s_count / s_base are rewritten by the next iteration
    }
}

// ------------------------------------------------------------------------------------------
// K4: exclusive scan of the chunk histogram — one CTA per 8192-entry tile, then one CTA over the
// (<= 4096) tile totals.  The rank kernel adds the two levels, so there is no third pass.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) ccl_scan_tiles_kernel(CclView v) {
    __shared__ int warp_sums[32];
    __shared__ int total;
    ccl_scan_tile_body(v, blockIdx.x, warp_sums, &total);
}

__global__ void __launch_bounds__(1024) ccl_scan_top_kernel(CclView v) {
    __shared__ int warp_sums[32];
    __shared__ int total;
    ccl_scan_top_body(v, warp_sums, &total);
}

// ------------------------------------------------------------------------------------------
// K5: raster-order rank of each global root -> label code
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ccl_rank_kernel(CclView v) {
    unsigned n = v.hdr->n_global_roots;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) ccl_rank_root(v, v.groots[i]);
}

// Planar stacks (2-D mode) have a component per object cross-section — 1.3 M roots on a 64 x 4096 x 4096 stack — and the
// per-root kernel above walks up to 63 words of the root bitmap for every one of them (0.17 ms, plus 0.04 ms to clear the
// bitmap afterwards).  This form streams the bitmap instead: a warp takes one 64-word chunk (two words per lane, one
// coalesced kilobyte), a warp scan of the popcounts gives every lane the rank of its first root, the label codes are
// written from the bit positions, and the words are zeroed on the way out (the clear kernel's job).  Needs whole chunks
// per plane (Y * ZW a multiple of 64), which also makes the per-plane restart of the numbering a chunk-table look-up.
__global__ void __launch_bounds__(256) ccl_rank_stream_kernel(CclView v) {
    const int lane = threadIdx.x & 31;
    const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long plane_chunks = ((long long)v.Y * v.ZW) >> 6;
    const int label0 = v.hdr->label_base + 1;
    for (long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < v.n_chunks; c += n_warps) {
        ull* words = v.rootbits + (c << 6) + 2 * lane;
        const uint4 q = *reinterpret_cast<const uint4*>(words);
        const ull w0 = ((ull)q.y << 32) | q.x, w1 = ((ull)q.w << 32) | q.z;
        if (!__any_sync(0xffffffffu, (w0 | w1) != 0ull)) continue;
        const int cnt = __popcll(w0) + __popcll(w1);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (cnt == 0) continue;  // lanes without roots are done (no further warp-wide operations below)
        int rank = v.chunks[c] + v.scan_tiles[c / SKB_SCAN_TILE] + incl - cnt;
        if (!v.connect_x) {  // planar: numbering restarts in every x-plane (= a whole number of chunks)
            const long long pc = (c / plane_chunks) * plane_chunks;
            rank -= v.chunks[pc] + v.scan_tiles[pc / SKB_SCAN_TILE];
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long wi = (c << 6) + 2 * lane + h;
            const long long rowi = wi / v.ZW;
            const int vbase = (int)(rowi * v.Z + (wi - rowi * v.ZW) * 64);
            for (ull m = h ? w1 : w0; m; m &= m - 1) v.parent[vbase + (__ffsll((long long)m) - 1)] = -(label0 + rank++);
        }
        *reinterpret_cast<uint4*>(words) = make_uint4(0u, 0u, 0u, 0u);  // leaves the bitmap clean for the next pass
    }
}

// leaves the root bitmap all-zero again (only the words the global roots touched), so the next pass over
// the same workspace can skip a V/8-byte memset
__global__ void __launch_bounds__(256) ccl_clear_rootbits_kernel(CclView v) {
    unsigned n = min(v.hdr->n_global_roots, (unsigned)v.capacity);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int bit;
        v.rootbits[word_of_voxel(v, v.groots[i], &bit)] = 0ull;
    }
}

// K6: tile roots that are not global roots take the code of their global root
__global__ void __launch_bounds__(256) ccl_publish_kernel(CclView v) {
    unsigned n = min(v.hdr->n_tile_roots, (unsigned)v.capacity);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int r = v.tile_roots[i], g = v.flat[i];
        if (g != r) v.parent[r] = v.parent[g];
    }
}

// ------------------------------------------------------------------------------------------
// dense writer: 8 voxels (one byte of the bit mask) per thread
// ------------------------------------------------------------------------------------------
// 8 voxels (one byte of the bit mask) per thread.  FLAT (Z % 64 == 0: no row padding, bit index == voxel index):
// the byte is bits[gi] and the voxels are 8 gi .. 8 gi + 7 — no division at all (the 64-bit div/mod per thread
// of the general form made this store-bound kernel run at half of HBM speed on a 64 x 4096 x 4096 stack).
template <typename OutT, bool FLAT>
__global__ void __launch_bounds__(256) ccl_dense_kernel(const ull* __restrict__ bits, const int* __restrict__ parent,
                                                       int Z, int ZW, int Z8, unsigned n_groups,
                                                       OutT* __restrict__ out, int vec_ok) {
    const unsigned gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_groups) return;
    unsigned byte;
    size_t vox;
    int z = 0;
    if (FLAT) {
        byte = __ldg(reinterpret_cast<const unsigned char*>(bits) + gi);
        vox = (size_t)gi * 8;
    } else {
        const unsigned rowi = gi / (unsigned)Z8;
        z = (int)(gi - rowi * (unsigned)Z8) * 8;
        byte = (unsigned)(bits[(size_t)rowi * ZW + (z >> 6)] >> (z & 63)) & 0xFFu;
        vox = (size_t)rowi * Z + z;
    }
    unsigned lab[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) lab[j] = 0u;
    if (byte) {
        // two dependent loads per foreground voxel, issued stage by stage for all 8 (index 0 stands in for background)
        int p1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) p1[j] = __ldg(parent + (((byte >> j) & 1u) ? vox + j : vox + (__ffs((int)byte) - 1)));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p2 = p1[j] < 0 ? p1[j] : __ldg(parent + p1[j]);
            lab[j] = ((byte >> j) & 1u) ? (unsigned)(-p2) : 0u;
        }
    }
    if (vec_ok) {
        if (sizeof(OutT) == 2) {
            skb_st_stream16(out + vox, make_uint4((lab[0] & 0xffffu) | (lab[1] << 16), (lab[2] & 0xffffu) | (lab[3] << 16),
                                                  (lab[4] & 0xffffu) | (lab[5] << 16), (lab[6] & 0xffffu) | (lab[7] << 16)));
        } else {
            skb_st_stream16(out + vox, make_uint4(lab[0], lab[1], lab[2], lab[3]));
            skb_st_stream16(out + vox + 4, make_uint4(lab[4], lab[5], lab[6], lab[7]));
        }
    } else {
        for (int j = 0; j < 8 && z + j < Z; ++j) out[vox + j] = (OutT)lab[j];
    }
}

// int32 labels, FLAT volumes: the same work with the lanes arranged so that EACH store instruction of a warp writes 512
// contiguous bytes.  In the kernel above a lane owns 8 consecutive voxels = 32 bytes and writes them with two 16-byte
// stores, so each instruction touches all 32 sectors of the warp's kilobyte but fills only half of each.  Here lane l owns
// the voxels [4l, 4l+4) of the warp's first 128 voxels and of its second 128 (two nibbles of the bit mask).
__global__ void __launch_bounds__(256) ccl_dense32_kernel(const unsigned char* __restrict__ bits, const int* __restrict__ parent,
                                                         unsigned n_groups, int* __restrict__ out) {
    const unsigned gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_groups) return;  // n_groups is a multiple of 32 (V % 256 == 0): whole warps leave together
    const unsigned lane = threadIdx.x & 31u, wbase = gi - lane;  // wbase = first byte of the warp's 32 bytes of bit mask
    const unsigned ba = __ldg(bits + wbase + (lane >> 1)), bb = __ldg(bits + wbase + 16u + (lane >> 1));
    const unsigned sh = (lane & 1u) * 4u;
    const unsigned nib[2] = {(ba >> sh) & 0xFu, (bb >> sh) & 0xFu};
    const size_t vox[2] = {(size_t)wbase * 8 + 4 * lane, (size_t)wbase * 8 + 128 + 4 * lane};
    unsigned lab[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) lab[h][j] = 0u;
    if (nib[0] | nib[1]) {
        int p1[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 4; ++j) p1[h][j] = ((nib[h] >> j) & 1u) ? __ldg(parent + vox[h] + j) : -1;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!((nib[h] >> j) & 1u)) continue;
                const int p2 = p1[h][j] < 0 ? p1[h][j] : __ldg(parent + p1[h][j]);
                lab[h][j] = (unsigned)(-p2);
            }
    }
    skb_st_stream16(out + vox[0], make_uint4(lab[0][0], lab[0][1], lab[0][2], lab[0][3]));
    skb_st_stream16(out + vox[1], make_uint4(lab[1][0], lab[1][1], lab[1][2], lab[1][3]));
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
CclView skb_ccl_make_view(const SkbCclLayout& L, void* ws, int planar, int64_t capacity, uint32_t* status, int32_t* ncomp) {
    char* base = static_cast<char*>(ws);
    CclView v;
    v.X = L.X; v.Y = L.Y; v.Z = L.Z; v.ZW = L.ZW;
    v.connect_x = planar ? 0 : 1;
    v.z_off = 0; v.Zl = L.Z; v.k0 = 0; v.nk = L.ZW;
    v.nk_shift = shift_of(L.ZW); v.y_shift = shift_of(L.Y);
    v.capacity = (int)capacity;
    v.hdr = reinterpret_cast<SkbCclHeader*>(base);
    v.bits = reinterpret_cast<ull*>(base + L.off_bits);
    v.parent = reinterpret_cast<int*>(base + L.off_parent);
    v.rootbits = reinterpret_cast<ull*>(base + L.off_rootbits);
    v.face_lo = nullptr; v.face_hi = nullptr;
    v.chunks = reinterpret_cast<int*>(base + L.off_chunks);
    v.cursors = reinterpret_cast<unsigned*>(base + L.off_cursors);
    v.scan_tiles = reinterpret_cast<int*>(base + L.off_scan_tiles);
    v.tile_roots = reinterpret_cast<int*>(base + L.off_tile_roots);
    v.flat = reinterpret_cast<int*>(base + L.off_flat);
    v.groots = reinterpret_cast<int*>(base + L.off_groots);
    v.status = status;
    v.ncomp_out = ncomp;
    v.n_words = L.n_words;
    v.n_chunks = L.n_chunks;
    v.n_scan_tiles = L.n_scan_tiles;
    return v;
}

extern "C" size_t skb_ccl_workspace_bytes(int64_t X, int64_t Y, int64_t Z, int64_t capacity) {
    if (X <= 0 || Y <= 0 || Z <= 0 || capacity <= 0) return 0;
    return skb_ccl_layout(X, Y, Z, capacity).total;
}

template <typename MaskT>
static void launch_pack(const void* mask, const CclView& v, cudaStream_t st) {
    const MaskT* m = static_cast<const MaskT*>(mask);
    const int nk_shift = shift_of(v.nk);
    const long long rows = (long long)v.X * v.Y;
    const bool fast = (v.Zl % 64 == 0) && skb_aligned16(mask);
    if (fast) {
        const unsigned n_seg = (unsigned)(rows * v.Zl / 16);
        ccl_pack_kernel<MaskT><<<(n_seg + 256 * PACK_SEGS - 1) / (256 * PACK_SEGS), 256, 0, st>>>(m, v, n_seg, nk_shift);
    } else {
        ccl_pack_generic_kernel<MaskT><<<(unsigned)((rows * v.nk + 255) / 256), 256, 0, st>>>(m, v);
    }
}

static void launch_tile(const CclView& v, int flags, cudaStream_t st) {
    const int nk_shift = shift_of(v.nk);
    const long long xt = (v.X + 7) / 8, n_yk = (long long)((v.Y + 7) / 8) * v.nk;
    const long long n_tiles = xt * n_yk;
    long long blocks = (n_tiles + CCL_TILE_WARPS - 1) / CCL_TILE_WARPS;
    if (blocks > 148 * 5) blocks = 148 * 5;  // 5 resident CTAs per SM (40 KB of shared memory each), persistent warps
    int dynamic = n_tiles >= 32 * blocks * CCL_TILE_WARPS ? 1 : 0;
    if (flags & SKB_CCL_TILES_DYNAMIC) dynamic = 1;
    if (flags & SKB_CCL_TILES_STATIC) dynamic = 0;
    ccl_tile_kernel<<<(unsigned)blocks, 32 * CCL_TILE_WARPS, 0, st>>>(v, (unsigned)n_tiles, (unsigned)n_yk, nk_shift,
                                                                    n_yk < (1LL << 30) ? shift_of((int)n_yk) : -1, dynamic);
}

// the part of a labelling pass selected by flags (SKB_CCL_PHASE_*): header + clears + pack, and/or the tile kernel
void skb_ccl_launch_pack_and_tile(const void* mask, int mask_dtype, const CclView& v, const SkbCclLayout& L,
                                  const SkbCclHeader& h, int flags, cudaStream_t st) {
    const bool all = !(flags & (SKB_CCL_PHASE_PACK | SKB_CCL_PHASE_LABEL));
    if (all || (flags & SKB_CCL_PHASE_PACK)) {
        char* base = reinterpret_cast<char*>(v.hdr);
        ccl_init_kernel<<<1, SKB_TILE_CURSORS, 0, st>>>(v, h, (flags & SKB_CCL_KEEP_STATUS) ? 1 : 0);  // header by value: no host->device copy on the path
        if (!(flags & SKB_CCL_WORKSPACE_CLEAN)) cudaMemsetAsync(base + L.off_rootbits, 0, (size_t)L.n_words * 8, st);
        cudaMemsetAsync(base + L.off_chunks, 0, (size_t)(L.n_chunks + 1) * 4, st);
        if (mask_dtype == SKB_U8) launch_pack<uint8_t>(mask, v, st);
        else launch_pack<int16_t>(mask, v, st);
    }
    if (all || (flags & SKB_CCL_PHASE_LABEL)) launch_tile(v, flags, st);
}

void skb_ccl_launch_boundary(const CclView& v, const SkbCclLayout& L, int TX, int TY, cudaStream_t st) {
    const int pair_mode = (v.n_words % 2 == 0 && (v.nk % 2 == 0 || v.nk == 1)) ? 1 : 0;
    const long long threads = pair_mode ? v.n_words / 2 : v.n_words;
    ccl_boundary_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(v, TX, TY, pair_mode);
}

extern "C" int skb_ccl_label_sparse(const void* mask, int mask_dtype, int64_t X, int64_t Y, int64_t Z, int planar,
                                    int32_t label_base, int64_t capacity, void* workspace, size_t workspace_bytes,
                                    int32_t* ncomp, uint32_t* status, int flags, void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_ccl_label_sparse");
    if (rc) return rc;
    SKB_REQUIRE(mask && workspace && status, "skb_ccl_label_sparse: NULL pointer");
    SKB_REQUIRE(mask_dtype == SKB_U8 || mask_dtype == SKB_I16, "skb_ccl_label_sparse: mask dtype must be u8 or i16");
    SKB_REQUIRE(label_base >= 0, "skb_ccl_label_sparse: label_base must be >= 0");
    SKB_REQUIRE(capacity > 0 && capacity <= 0x7fffffff, "skb_ccl_label_sparse: bad capacity");
    SKB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "skb_ccl_label_sparse: workspace must be 256-byte aligned");
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, capacity);
    if (workspace_bytes < L.total) {
        skb_set_error("skb_ccl_label_sparse: workspace %zu < required %zu bytes", workspace_bytes, L.total);
        return SKB_E_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CclView v = skb_ccl_make_view(L, workspace, planar, capacity, status, ncomp);

    SkbCclHeader h = {};
    h.label_base = label_base; h.planar = planar; h.capacity = (int)capacity;
    h.dims[0] = L.X; h.dims[1] = L.Y; h.dims[2] = L.Z;
    skb_ccl_launch_pack_and_tile(mask, mask_dtype, v, L, h, flags, st);
    SKB_LAUNCH_CHECK("ccl_pack/tile_kernel");
    if ((flags & SKB_CCL_PHASE_PACK) && !(flags & SKB_CCL_PHASE_LABEL)) return SKB_OK;

    const int TY = 8, TX = 8;
    skb_ccl_launch_boundary(v, L, TX, TY, st);
    const int list_grid = 148 * 4;
    ccl_flatten_kernel<<<list_grid, 256, 0, st>>>(v);
    ccl_scan_tiles_kernel<<<(unsigned)L.n_scan_tiles, 1024, 0, st>>>(v);
    ccl_scan_top_kernel<<<1, 1024, 0, st>>>(v);
    if (planar && ((long long)L.Y * L.ZW) % 64 == 0 && L.n_words % 64 == 0) {
        ccl_rank_stream_kernel<<<148 * 8, 256, 0, st>>>(v);  // many small components: stream the root bitmap (and clear it)
    } else {
        ccl_rank_kernel<<<list_grid, 256, 0, st>>>(v);
        ccl_clear_rootbits_kernel<<<list_grid, 256, 0, st>>>(v);
    }
    ccl_publish_kernel<<<list_grid, 256, 0, st>>>(v);
    SKB_LAUNCH_CHECK("ccl merge kernels");
    return SKB_OK;
}

extern "C" int skb_ccl_write_dense(const void* workspace, int64_t X, int64_t Y, int64_t Z, void* out, int out_dtype,
                                   void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_ccl_write_dense");
    if (rc) return rc;
    SKB_REQUIRE(workspace && out, "skb_ccl_write_dense: NULL pointer");
    SKB_REQUIRE(out_dtype == SKB_I16 || out_dtype == SKB_I32, "skb_ccl_write_dense: out dtype must be i16 or i32");
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, 1);
    const char* base = static_cast<const char*>(workspace);
    const ull* bits = reinterpret_cast<const ull*>(base + L.off_bits);
    const int* parent = reinterpret_cast<const int*>(base + L.off_parent);
    int Z8 = (int)((Z + 7) / 8);
    const unsigned groups = (unsigned)((long long)X * Y * Z8);  // <= 2^28
    unsigned nb = (groups + 255u) / 256u;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int vec_ok = (Z % 8 == 0) && skb_aligned16(out) ? 1 : 0;
    const bool flat = Z % 64 == 0 && vec_ok;
    // (a persistent form — one 128-byte load of bits per warp and step, shuffled bytes, next word prefetched — was
    //  measured on 64 x 4096 x 4096 and was 5 % SLOWER than this one-byte-per-thread form: 2.83 vs 2.69 ms for the labelling
    //  + dense write; the kernel is bound by its 4 B/voxel of stores either way)
    if (out_dtype == SKB_I16) {
        if (flat) ccl_dense_kernel<int16_t, true><<<nb, 256, 0, st>>>(bits, parent, (int)Z, L.ZW, Z8, groups, static_cast<int16_t*>(out), vec_ok);
        else ccl_dense_kernel<int16_t, false><<<nb, 256, 0, st>>>(bits, parent, (int)Z, L.ZW, Z8, groups, static_cast<int16_t*>(out), vec_ok);
    } else {
        if (flat && (X * Y * Z) % 256 == 0 && !getenv("SKB_DENSE_OLD"))
            ccl_dense32_kernel<<<nb, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(bits), parent, groups, static_cast<int32_t*>(out));
        else if (flat) ccl_dense_kernel<int32_t, true><<<nb, 256, 0, st>>>(bits, parent, (int)Z, L.ZW, Z8, groups, static_cast<int32_t*>(out), vec_ok);
        else ccl_dense_kernel<int32_t, false><<<nb, 256, 0, st>>>(bits, parent, (int)Z, L.ZW, Z8, groups, static_cast<int32_t*>(out), vec_ok);
    }
    SKB_LAUNCH_CHECK("ccl_dense_kernel");
    return SKB_OK;
}

// ---- launchers used by skb_shard.cu -------------------------------------------------------------------
void skb_ccl_launch_scan_and_rank(const CclView& v, const SkbCclLayout& L, cudaStream_t st) {
    ccl_scan_tiles_kernel<<<(unsigned)L.n_scan_tiles, 1024, 0, st>>>(v);
    ccl_scan_top_kernel<<<1, 1024, 0, st>>>(v);
    ccl_rank_kernel<<<148 * 4, 256, 0, st>>>(v);
}

void skb_ccl_launch_publish(const CclView& v, cudaStream_t st) { ccl_publish_kernel<<<148 * 4, 256, 0, st>>>(v); }

