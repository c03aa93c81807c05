// Connected-component labelling of the thresholded skeleton mask on B200 (sm_100a).
//
// Replaces the arithmetic of skoots/lib/flood_fill.py:13-140 (scipy.ndimage.label per crop +
// seam merging).  Design:
//
//   K1 ccl_tile_kernel      one 8x8x64 tile per 64-thread CTA, one (x,y) row per thread.  The mask
//                           is read once with 16-byte loads (Z is the contiguous axis), packed
//                           into one 64-bit word per row, written out as the bit-packed mask,
//                           and labelled inside the tile by a shared-memory union-find over
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

#include "skb_common.cuh"

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
struct CclView {
    int X, Y, Z, ZW;
    int connect_x;  // 0 in planar (per-x-plane, 4-connectivity) mode
    int ZW_tiles;   // z tiles per row in the tile kernel
    int capacity;
    SkbCclHeader* hdr;
    ull* bits;
    int* parent;
    ull* rootbits;
    int* chunks;
    int* scan_tiles;
    int* tile_roots;
    int* flat;
    int* groots;
    unsigned* status;
    int* ncomp_out;
    long long n_words, n_chunks, n_scan_tiles;
};

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

__device__ __forceinline__ int gload(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

__device__ __forceinline__ int gfind(const int* parent, int a) {
    int p = gload(parent + a);
    while (p != a) {
        a = p;
        p = gload(parent + a);
    }
    return a;
}

__device__ __forceinline__ void gunion(int* parent, int a, int b) {
    for (;;) {
        a = gfind(parent, a);
        b = gfind(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(parent + a, b);
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

// foreground bits of the n (<= TZ) mask elements of one row segment
template <typename MaskT, int TZ>
__device__ __forceinline__ ull row_bits(const MaskT* p, int n, bool vec_ok);

template <typename MaskT, int TZ>
__device__ __forceinline__ ull row_bits_scalar(const MaskT* p, int n) {
    ull w = 0;
    for (int i = 0; i < n; ++i) w |= (ull)(p[i] > 0) << i;
    return w;
}

template <int TZ>
__device__ __forceinline__ ull row_bits_u8(const uint8_t* p, int n, bool vec_ok) {
    if (!vec_ok) return row_bits_scalar<uint8_t, TZ>(p, n);
    ull w = 0;
    uint4 q[TZ / 16];
#pragma unroll
    for (int k = 0; k < TZ / 16; ++k)  // all loads first; groups past the end of the row read nothing
        q[k] = (16 * (k + 1) <= n) ? __ldg(reinterpret_cast<const uint4*>(p) + k) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < TZ / 16; ++k) {
        unsigned h = nz4(q[k].x) | (nz4(q[k].y) << 4) | (nz4(q[k].z) << 8) | (nz4(q[k].w) << 12);
        w |= (ull)h << (16 * k);
    }
    for (int i = n & ~15; i < n; ++i) w |= (ull)(p[i] != 0) << i;
    return w;
}

template <int TZ>
__device__ __forceinline__ ull row_bits_i16(const int16_t* p, int n, bool vec_ok) {
    if (!vec_ok) return row_bits_scalar<int16_t, TZ>(p, n);
    ull w = 0;
#pragma unroll
    for (int k = 0; k < TZ / 8; ++k) {
        if (8 * (k + 1) <= n) {
            uint4 q = __ldg(reinterpret_cast<const uint4*>(p) + k);
            unsigned h = gt2(q.x) | (gt2(q.y) << 2) | (gt2(q.z) << 4) | (gt2(q.w) << 6);
            w |= (ull)h << (8 * k);
        }
    }
    for (int i = n & ~7; i < n; ++i) w |= (ull)(p[i] > 0) << i;
    return w;
}

template <typename MaskT, int TZ> struct RowBits;
template <int TZ> struct RowBits<uint8_t, TZ> {
    static __device__ __forceinline__ ull get(const uint8_t* p, int n, bool v) { return row_bits_u8<TZ>(p, n, v); }
};
template <int TZ> struct RowBits<int16_t, TZ> {
    static __device__ __forceinline__ ull get(const int16_t* p, int n, bool v) { return row_bits_i16<TZ>(p, n, v); }
};

__global__ void ccl_init_kernel(CclView v, SkbCclHeader h) {
    if (threadIdx.x == 0) {
        *v.hdr = h;
        *v.status = 0u;
        if (v.ncomp_out) *v.ncomp_out = 0;
    }
}

// ------------------------------------------------------------------------------------------
// K1: tile-local run-based union-find
// ------------------------------------------------------------------------------------------
// One CTA = one 8 x 8 x TZ tile, one THREAD = one (x,y) row of it: the thread loads its whole
// TZ-element row segment (4 independent 16-byte loads for u8, TZ = 64), packs it into a 64-bit
// word, and does every union of that row with its y-1 / x-1 neighbour rows.  A run start can only
// sit at every other bit, so the union-find array needs 32 slots per row: slot = row*32 + (p>>1).
template <typename MaskT, int TZ>
__global__ void __launch_bounds__(64) ccl_tile_kernel(const MaskT* __restrict__ mask, CclView v, int vec_ok) {
    constexpr int TY = 8;
    __shared__ ull srow[64];
    __shared__ int slab[64 * 32];
    __shared__ int s_count, s_base;

    const int row = threadIdx.x;
    const int ly = row & 7, lx = row >> 3;
    // linear CTA index -> (z tile fastest, then y, then x); avoids the 65535 limit of grid.y/z
    const unsigned nzt = (unsigned)v.ZW_tiles, nyt = (unsigned)((v.Y + 7) >> 3);
    const unsigned bq = blockIdx.x / nzt, zt = blockIdx.x - bq * nzt;
    const unsigned xt = bq / nyt, yt = bq - xt * nyt;
    const int x0 = (int)xt * 8, y0 = (int)yt * 8, z0 = (int)zt * TZ;
    const int x = x0 + lx, y = y0 + ly;
    const bool in_row = (x < v.X) && (y < v.Y);
    const unsigned rowi = (unsigned)x * (unsigned)v.Y + (unsigned)y;

    if (row == 0) s_count = 0;
    ull w = 0;
    if (in_row) {
        w = RowBits<MaskT, TZ>::get(mask + (size_t)rowi * v.Z + z0, min(TZ, v.Z - z0), vec_ok != 0);
        v.bits[(size_t)rowi * v.ZW + zt] = w;
    }
    srow[row] = w;
    if (!__syncthreads_or(w != 0ull)) return;  // empty tile: nothing to label

    const ull starts = w & ~(w << 1);
    for (ull s = starts; s; s &= s - 1) {
        int p = __ffsll((long long)s) - 1;
        slab[row * 32 + (p >> 1)] = row * 32 + (p >> 1);
    }
    __syncthreads();

    // unions with the y-1 and x-1 rows of the same tile: one per maximal joint run
    if (w) {
        if (ly > 0) {
            ull wn = srow[row - 1], a = w & wn;
            for (ull s = a & ~(a << 1); s; s &= s - 1) {
                int p = __ffsll((long long)s) - 1;
                sunion(slab, row * 32 + (run_start(w, p) >> 1), (row - 1) * 32 + (run_start(wn, p) >> 1));
            }
        }
        if (lx > 0 && v.connect_x) {
            ull wn = srow[row - TY], a = w & wn;
            for (ull s = a & ~(a << 1); s; s &= s - 1) {
                int p = __ffsll((long long)s) - 1;
                sunion(slab, row * 32 + (run_start(w, p) >> 1), (row - TY) * 32 + (run_start(wn, p) >> 1));
            }
        }
    }
    __syncthreads();

    // resolve every run of my row; write parent for all its voxels
    ull rootmask = 0;  // bit p set when the run starting at p is a tile root
    const int gbase = (int)(rowi * (unsigned)v.Z) + z0;  // voxel index of bit 0 of this row word
    for (ull s = starts; s; s &= s - 1) {
        int p = __ffsll((long long)s) - 1;
        int l = row * 32 + (p >> 1);
        int r = sfind(slab, l);
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
        ull tt = ~(w >> p);
        int len = tt ? __ffsll((long long)tt) - 1 : 64 - p;
        for (int j = 0; j < len; ++j) v.parent[gbase + p + j] = groot;
    }
    int mine = __popcll(rootmask);
    int off = mine ? atomicAdd(&s_count, mine) : 0;
    __syncthreads();
    if (row == 0) s_base = (int)atomicAdd(&v.hdr->n_tile_roots, (unsigned)s_count);
    __syncthreads();
    if (mine) {
        int at = s_base + off;
        for (ull m = rootmask; m; m &= m - 1) {
            int p = __ffsll((long long)m) - 1;
            if (at < v.capacity) v.tile_roots[at] = gbase + p;
            else atomicOr(v.status, SKB_STATUS_ROOT_OVERFLOW);
            ++at;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2: unions across tile faces, driven by the bit-packed mask
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ccl_boundary_kernel(CclView v, int TX, int TY) {
    // n_words <= 2^31: all index math in 32 bits (64-bit div/mod costs ~100 instructions each)
    const unsigned widx = blockIdx.x * blockDim.x + threadIdx.x;
    if (widx >= (unsigned)v.n_words) return;
    const ull w = v.bits[widx];
    if (!w) return;
    const unsigned uzw = (unsigned)v.ZW, uyy = (unsigned)v.Y;
    const unsigned rowi = widx / uzw, k = widx - rowi * uzw;
    const unsigned x = rowi / uyy, y = rowi - x * uyy;
    const int gbase = (int)(rowi * (unsigned)v.Z + 64u * k);
    if (k > 0 && (w & 1ull)) {
        if (v.bits[widx - 1] >> 63) gunion(v.parent, gbase, gbase - 1);
    }
    if (y > 0 && (y % (unsigned)TY) == 0) {
        ull a = w & v.bits[widx - uzw];
        for (ull s = a & ~(a << 1); s; s &= s - 1) {
            int p = __ffsll((long long)s) - 1;
            gunion(v.parent, gbase + p, gbase + p - v.Z);
        }
    }
    if (v.connect_x && x > 0 && (x % (unsigned)TX) == 0) {
        ull a = w & v.bits[widx - uyy * uzw];
        int plane = v.Y * v.Z;
        for (ull s = a & ~(a << 1); s; s &= s - 1) {
            int p = __ffsll((long long)s) - 1;
            gunion(v.parent, gbase + p, gbase + p - plane);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K3: pointer jumping of tile roots; global roots -> bitmap + chunk histogram + list
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ long long word_of_voxel(const CclView& v, int vox, int* bit) {
    int zrow = vox % v.Z;
    long long rowi = vox / v.Z;
    *bit = zrow & 63;
    return rowi * v.ZW + (zrow >> 6);
}

__global__ void __launch_bounds__(256) ccl_flatten_kernel(CclView v) {
    unsigned n = min(v.hdr->n_tile_roots, (unsigned)v.capacity);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int r = v.tile_roots[i];
        int g = gfind(v.parent, r);
        v.flat[i] = g;
        if (g == r) {
            int bit;
            long long wi = word_of_voxel(v, r, &bit);
            atomicOr(&v.rootbits[wi], 1ull << bit);
            atomicAdd(&v.chunks[wi >> 6], 1);
            // warp-aggregated append: one atomic per warp, slots handed out by ballot rank
            unsigned m = __activemask();
            int lane = threadIdx.x & 31, leader = __ffs((int)m) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(&v.hdr->n_global_roots, (unsigned)__popc(m));
            base = __shfl_sync(m, base, leader);
            v.groots[base + __popc(m & ((1u << lane) - 1u))] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4: exclusive scan of the chunk histogram — one CTA per 8192-entry tile, then one CTA over the
// (<= 4096) tile totals.  The rank kernel adds the two levels, so there is no third pass.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_exclusive_scan_1024(int sum, int* warp_sums, int* total) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int ws = warp_sums[lane], wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_sums[lane] = wi - ws;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    return warp_sums[wid] + incl - sum;
}

__global__ void __launch_bounds__(1024) ccl_scan_tiles_kernel(CclView v) {
    __shared__ int warp_sums[32];
    __shared__ int total;
    constexpr int PER = SKB_SCAN_TILE / 1024;
    const long long at = (long long)blockIdx.x * SKB_SCAN_TILE + (long long)threadIdx.x * PER;
    int vals[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        vals[j] = (at + j < v.n_chunks) ? v.chunks[at + j] : 0;
        sum += vals[j];
    }
    int run = block_exclusive_scan_1024(sum, warp_sums, &total);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        if (at + j < v.n_chunks) v.chunks[at + j] = run;
        run += vals[j];
    }
    if (threadIdx.x == 0) v.scan_tiles[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) ccl_scan_top_kernel(CclView v) {
    __shared__ int warp_sums[32];
    __shared__ int total;
    constexpr int PER = 4;  // n_scan_tiles <= 4096 for any volume the library accepts
    const long long at = (long long)threadIdx.x * PER;
    int vals[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        vals[j] = (at + j < v.n_scan_tiles) ? v.scan_tiles[at + j] : 0;
        sum += vals[j];
    }
    int run = block_exclusive_scan_1024(sum, warp_sums, &total);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        if (at + j < v.n_scan_tiles) v.scan_tiles[at + j] = run;
        run += vals[j];
    }
    if (threadIdx.x == 0) {
        v.hdr->n_components = total;
        if (v.ncomp_out) *v.ncomp_out = total;
    }
}

// ------------------------------------------------------------------------------------------
// K5: raster-order rank of each global root -> label code
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int roots_before_word(const CclView& v, long long wi) {
    long long c = wi >> 6;
    int r = v.chunks[c] + v.scan_tiles[c / SKB_SCAN_TILE];
    for (long long j = c << 6; j < wi; ++j) r += __popcll(v.rootbits[j]);
    return r;
}

__global__ void __launch_bounds__(256) ccl_rank_kernel(CclView v) {
    unsigned n = v.hdr->n_global_roots;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int r = v.groots[i];
        int bit;
        long long wi = word_of_voxel(v, r, &bit);
        int rank = roots_before_word(v, wi) + __popcll(v.rootbits[wi] & ((1ull << bit) - 1ull));
        if (!v.connect_x) {  // planar: numbering restarts in every x-plane
            long long plane_words = (long long)v.Y * v.ZW;
            rank -= roots_before_word(v, (wi / plane_words) * plane_words);
        }
        v.parent[r] = -(v.hdr->label_base + 1 + rank);
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
template <typename OutT>
__global__ void __launch_bounds__(256) ccl_dense_kernel(const ull* __restrict__ bits, const int* __restrict__ parent,
                                                       int Z, int ZW, int Z8, long long n_groups,
                                                       OutT* __restrict__ out, int vec_ok) {
    long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_groups) return;
    long long rowi = gi / Z8;
    int z = (int)(gi % Z8) * 8;
    unsigned byte = (unsigned)(bits[rowi * ZW + (z >> 6)] >> (z & 63)) & 0xFFu;
    long long vox = rowi * Z + z;
    __align__(16) OutT lab[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) lab[j] = (byte >> j) & 1u ? (OutT)skb_sparse_label(parent, (int)(vox + j)) : (OutT)0;
    if (vec_ok) {
        if (sizeof(OutT) == 2) {
            *reinterpret_cast<uint4*>(out + vox) = *reinterpret_cast<uint4*>(lab);
        } else {
            reinterpret_cast<uint4*>(out + vox)[0] = reinterpret_cast<uint4*>(lab)[0];
            reinterpret_cast<uint4*>(out + vox)[1] = reinterpret_cast<uint4*>(lab)[1];
        }
    } else {
        for (int j = 0; j < 8 && z + j < Z; ++j) out[vox + j] = lab[j];
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static CclView make_view(const SkbCclLayout& L, void* ws, int planar, int64_t capacity, uint32_t* status,
                         int32_t* ncomp) {
    char* base = static_cast<char*>(ws);
    CclView v;
    v.X = L.X; v.Y = L.Y; v.Z = L.Z; v.ZW = L.ZW;
    v.connect_x = planar ? 0 : 1;
    v.ZW_tiles = 1;
    v.capacity = (int)capacity;
    v.hdr = reinterpret_cast<SkbCclHeader*>(base);
    v.bits = reinterpret_cast<ull*>(base + L.off_bits);
    v.parent = reinterpret_cast<int*>(base + L.off_parent);
    v.rootbits = reinterpret_cast<ull*>(base + L.off_rootbits);
    v.chunks = reinterpret_cast<int*>(base + L.off_chunks);
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
static void launch_tile(const void* mask, const CclView& v, int tz, int vec_ok, cudaStream_t st) {
    const MaskT* m = static_cast<const MaskT*>(mask);
    CclView vv = v;
    vv.ZW_tiles = (v.Z + tz - 1) / tz;
    const unsigned grid = (unsigned)((long long)vv.ZW_tiles * ((v.Y + 7) / 8) * ((v.X + 7) / 8));
    if (tz == 64) ccl_tile_kernel<MaskT, 64><<<grid, 64, 0, st>>>(m, vv, vec_ok);
    else if (tz == 32) ccl_tile_kernel<MaskT, 32><<<grid, 64, 0, st>>>(m, vv, vec_ok);
    else ccl_tile_kernel<MaskT, 16><<<grid, 64, 0, st>>>(m, vv, vec_ok);
}

extern "C" int skb_ccl_label_sparse(const void* mask, int mask_dtype, int64_t X, int64_t Y, int64_t Z, int planar,
                                    int32_t label_base, int64_t capacity, void* workspace, size_t workspace_bytes,
                                    int32_t* ncomp, uint32_t* status, void* stream) {
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
    CclView v = make_view(L, workspace, planar, capacity, status, ncomp);

    SkbCclHeader h = {};
    h.label_base = label_base; h.planar = planar; h.capacity = (int)capacity;
    h.dims[0] = L.X; h.dims[1] = L.Y; h.dims[2] = L.Z;
    char* base = static_cast<char*>(workspace);
    ccl_init_kernel<<<1, 32, 0, st>>>(v, h);  // header by value: no host->device copy on the path
    cudaMemsetAsync(base + L.off_rootbits, 0, (size_t)L.n_words * 8, st);
    cudaMemsetAsync(base + L.off_chunks, 0, (size_t)(L.n_chunks + 1) * 4, st);

    const int tz = Z > 32 ? 64 : (Z > 16 ? 32 : 16);
    const int elem = mask_dtype == SKB_U8 ? 1 : 2;
    const int vec_ok = ((Z * elem) % 16 == 0) && skb_aligned16(mask) ? 1 : 0;
    if (mask_dtype == SKB_U8) launch_tile<uint8_t>(mask, v, tz, vec_ok, st);
    else launch_tile<int16_t>(mask, v, tz, vec_ok, st);
    SKB_LAUNCH_CHECK("ccl_tile_kernel");

    const int TY = 8, TX = 8;
    unsigned nb = (unsigned)((L.n_words + 255) / 256);
    ccl_boundary_kernel<<<nb, 256, 0, st>>>(v, TX, TY);
    const int list_grid = 148 * 4;
    ccl_flatten_kernel<<<list_grid, 256, 0, st>>>(v);
    ccl_scan_tiles_kernel<<<(unsigned)L.n_scan_tiles, 1024, 0, st>>>(v);
    ccl_scan_top_kernel<<<1, 1024, 0, st>>>(v);
    ccl_rank_kernel<<<list_grid, 256, 0, st>>>(v);
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
    long long groups = (long long)X * Y * Z8;
    unsigned nb = (unsigned)((groups + 255) / 256);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int vec_ok = (Z % 8 == 0) && skb_aligned16(out) ? 1 : 0;
    if (out_dtype == SKB_I16)
        ccl_dense_kernel<int16_t><<<nb, 256, 0, st>>>(bits, parent, (int)Z, L.ZW, Z8, groups, static_cast<int16_t*>(out), vec_ok);
    else
        ccl_dense_kernel<int32_t><<<nb, 256, 0, st>>>(bits, parent, (int)Z, L.ZW, Z8, groups, static_cast<int32_t*>(out), vec_ok);
    SKB_LAUNCH_CHECK("ccl_dense_kernel");
    return SKB_OK;
}
