// Shared between skb_ccl.cu (single-GPU labelling) and skb_shard.cu (Z-sharded labelling): the view of a CCL
// workspace the kernels take by value, the global-memory union-find helpers, and the host-side launchers that
// skb_shard.cu needs from skb_ccl.cu (kernels cannot be launched across translation units without -rdc, plain
// host functions can be called).
#pragma once
#include "skb_common.cuh"

struct CclView {
    int X, Y, Z, ZW;
    int connect_x;  // 0 in planar (per-x-plane, 4-connectivity) mode
    int z_off, Zl;  // the slab [z_off, z_off+Zl) of the global volume held by `mask` (whole volume: 0, Z)
    int k0, nk;     // the slab covers words [k0, k0+nk) of every row; `bits` holds exactly those: nk words per row
    int nk_shift, y_shift;  // log2(nk), log2(Y) when they are powers of two, else -1
    int capacity;
    SkbCclHeader* hdr;
    ull* bits;
    int* parent;
    ull* rootbits;
    ull* face_lo;   // sharded mode only (else NULL): word k0 / k0+nk-1 of every row, contiguous
    ull* face_hi;
    int* chunks;
    unsigned* cursors;  // SKB_TILE_CURSORS work cursors, SKB_TILE_CURSOR_STRIDE ints apart
    int* scan_tiles;
    int* tile_roots;
    int* flat;
    int* groots;
    unsigned* status;
    int* ncomp_out;
    long long n_words, n_chunks, n_scan_tiles;
};

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

__device__ __forceinline__ long long word_of_voxel(const CclView& v, int vox, int* bit) {
    int zrow = vox % v.Z;
    long long rowi = vox / v.Z;
    *bit = zrow & 63;
    return rowi * v.ZW + (zrow >> 6);
}

// ---- numbering: exclusive scan of the chunk histogram (one 1024-thread CTA per SKB_SCAN_TILE-entry tile, then one CTA
// over the <= 4096 tile totals; the rank adds the two levels, so there is no third pass) and the raster rank of a root.
// Device functions so that both the stand-alone kernels (skb_ccl.cu) and the fused merge kernel (skb_shard.cu) use them.
__device__ __forceinline__ int block_exclusive_scan_1024(int sum, int* warp_sums, int* total) {  // blockDim.x == 1024
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

__device__ __forceinline__ void ccl_scan_tile_body(const CclView& v, long long tile, int* warp_sums, int* total) {
    constexpr int PER = SKB_SCAN_TILE / 1024;
    const long long at = tile * SKB_SCAN_TILE + (long long)threadIdx.x * PER;
    int vals[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        vals[j] = (at + j < v.n_chunks) ? v.chunks[at + j] : 0;
        sum += vals[j];
    }
    int run = block_exclusive_scan_1024(sum, warp_sums, total);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        if (at + j < v.n_chunks) v.chunks[at + j] = run;
        run += vals[j];
    }
    if (threadIdx.x == 0) v.scan_tiles[tile] = *total;
}

__device__ __forceinline__ void ccl_scan_top_body(const CclView& v, int* warp_sums, int* total) {
    constexpr int PER = 4;  // n_scan_tiles <= 4096 for any volume the library accepts
    const long long at = (long long)threadIdx.x * PER;
    int vals[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        vals[j] = (at + j < v.n_scan_tiles) ? v.scan_tiles[at + j] : 0;
        sum += vals[j];
    }
    int run = block_exclusive_scan_1024(sum, warp_sums, total);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        if (at + j < v.n_scan_tiles) v.scan_tiles[at + j] = run;
        run += vals[j];
    }
    if (threadIdx.x == 0) {
        v.hdr->n_components = *total;
        if (v.ncomp_out) *v.ncomp_out = *total;
    }
}

__device__ __forceinline__ int roots_before_word(const CclView& v, long long wi) {
    long long c = wi >> 6;
    int r = v.chunks[c] + v.scan_tiles[c / SKB_SCAN_TILE];
    for (long long j = c << 6; j < wi; ++j) r += __popcll(v.rootbits[j]);
    return r;
}

// label code of global root r: -(label_base + 1 + raster rank)
__device__ __forceinline__ void ccl_rank_root(const CclView& v, int r) {
    int bit;
    long long wi = word_of_voxel(v, r, &bit);
    int rank = roots_before_word(v, wi) + __popcll(v.rootbits[wi] & ((1ull << bit) - 1ull));
    if (!v.connect_x) {  // planar: numbering restarts in every x-plane
        long long plane_words = (long long)v.Y * v.ZW;
        rank -= roots_before_word(v, (wi / plane_words) * plane_words);
    }
    v.parent[r] = -(v.hdr->label_base + 1 + rank);
}

// log2 for powers of two, else -1
static inline int shift_of(int n) {
    if (n <= 0 || (n & (n - 1))) return -1;
    int s = 0;
    while ((1 << s) < n) ++s;
    return s;
}

// ---- host side, defined in skb_ccl.cu ---------------------------------------------------------------
CclView skb_ccl_make_view(const SkbCclLayout& L, void* ws, int planar, int64_t capacity, uint32_t* status, int32_t* ncomp);
// the part of a labelling pass selected by flags (SKB_CCL_PHASE_*): header + clears + pack, and/or the tile kernel
void skb_ccl_launch_pack_and_tile(const void* mask, int mask_dtype, const CclView& v, const SkbCclLayout& L,
                                  const SkbCclHeader& h, int flags, cudaStream_t st);
void skb_ccl_launch_boundary(const CclView& v, const SkbCclLayout& L, int TX, int TY, cudaStream_t st);
// chunk histogram -> raster rank of every listed global root -> label codes in parent[root]
void skb_ccl_launch_scan_and_rank(const CclView& v, const SkbCclLayout& L, cudaStream_t st);
// tile roots that are not global roots take the code of their global root
void skb_ccl_launch_publish(const CclView& v, cudaStream_t st);
