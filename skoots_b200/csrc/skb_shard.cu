// Z-sharded connected-component labelling and its two exchanges: the kernels a rank
// runs around the single-GPU labelling kernels of skb_ccl.cu, for both transports (NCCL between the phases, or
// stores into peer mailboxes over NVLink with release/acquire flags).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "skb_ccl.cuh"

namespace cg = cooperative_groups;

// ==========================================================================================
// Z-sharded labelling.
//
// Every rank owns the slab z in [z_off, z_off+Zl) (z_off, Zl multiples of 64) of the mask, but its
// union-find lives in the GLOBAL voxel index space, so component ids (= the smallest global voxel
// index of the component) mean the same thing on every rank and no id translation is ever needed.
//
//   local   : init + tile kernel + boundary kernel on the slab, pointer-jump the tile roots,
//             list the slab's roots                                    (skb_shard_label_local)
//   runs    : the z-runs inside the H boundary planes of the slab, each with its root id
//             -> sent to the Z-neighbour (NCCL send/recv)               (skb_shard_emit_runs)
//   ingest  : the neighbour's runs become my halo: one 64-bit word per row (bits) and
//             parent[halo voxel] = the neighbour's root id             (skb_shard_ingest_runs)
//   pairs   : (my root, neighbour root) wherever my last plane touches the halo's first plane;
//             packed behind my root list -> all-gathered (NCCL)        (skb_shard_boundary_pairs)
//   merge   : every rank applies every pair to its own union-find, ranks all global roots in
//             raster order (identical numbering on every rank = the single-GPU numbering) and
//             publishes the label codes                                 (skb_shard_merge)
// ==========================================================================================
static const ull* face_words(const SkbCclLayout& L, const void* workspace, int high, int64_t Zl);
static int shard_common(const char* who, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl);

__global__ void __launch_bounds__(256) shard_local_roots_kernel(CclView v) {
    unsigned n = min(v.hdr->n_tile_roots, (unsigned)v.capacity);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int r = v.tile_roots[i];
        int g = gfind(v.parent, r);
        v.flat[i] = g;
        if (g == r) {
            unsigned m = __activemask();
            int lane = threadIdx.x & 31, leader = __ffs((int)m) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(&v.hdr->n_global_roots, (unsigned)__popc(m));
            base = __shfl_sync(m, base, leader);
            v.groots[base + __popc(m & ((1u << lane) - 1u))] = r;
        } else {
            // path compression: every voxel is now two hops from its slab root (voxel -> tile root -> root), which is
            // what the run emission and the face pairing chase.  Safe next to concurrent finds: a reader sees the old
            // parent or the root, both ancestors.
            v.parent[r] = g;
        }
    }
}

// ---- cross-GPU flags (peer transport): release/acquire at system scope ---------------------------
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// spins until *flag has reached pass `epoch` (flags only ever grow; the difference is wrap-safe).
// Bounded: ~8 s of SM clocks, then the pass carries on with SKB_STATUS_PEER_TIMEOUT set.
__device__ __forceinline__ void spin_until(const int* flag, int epoch, unsigned* status) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) - epoch < 0) {
        __nanosleep(64);
        if (clock64() - t0 > (16LL << 30)) {
            if (status) atomicOr(status, SKB_STATUS_PEER_TIMEOUT);
            break;
        }
    }
}

// where a face's runs go: `counter` is always local; `triples` is local (NCCL transport: runs + 3)
// or the neighbour's receive buffer (peer transport; the copy used is picked by the pass number)
struct RunsDst {
    int* counter;
    int* triples;
    const int* epoch;        // NULL: no parity
    long long parity_stride;
};

// one face of the slab: the planes [z_lo, z_hi) inside the face word, and where its runs go
struct EmitFace {
    const ull* face;  // compact copy of that word of every row
    int z_lo, z_hi;
    RunsDst dst;
    // peer transport: the LAST CTA of the face to finish publishes the run count in the neighbour's buffer and releases
    // its flag (round 1 used a second, one-warp kernel for this: one more launch on the latency chain).  NULL: no signal.
    int* done;         // local counter of finished CTAs (zero between passes)
    int* remote_runs;  // the neighbour's receive buffer, copy 0 ([count,_,_] + triples)
    int* remote_flag;
};

// counter = count, then (start voxel, length, root id) triples.  blockIdx.y picks the face (the peer
// transport emits both faces of a slab with one launch).  A thread takes EMIT_ROWS consecutive rows (two
// 16-byte loads of the compact face words): 99 % of the rows have nothing on a face, so the kernel is a
// stream with a vote, and the per-warp overhead is paid once per 128 rows.
constexpr int EMIT_ROWS = 4;
constexpr int EMIT_QUEUE = 96;  // runs a warp can queue

// returns whether this thread stored any triple
__device__ __forceinline__ bool emit_face_rows(const CclView& v, const EmitFace& f, int cap, unsigned* status, int2 (*s_runs)[EMIT_QUEUE],
                                               unsigned row_block) {
    const RunsDst& dst = f.dst;
    const int z_lo = f.z_lo, z_hi = f.z_hi;
    int* const runs = dst.counter;
    int* const tri = dst.triples + (dst.epoch ? (long long)(*dst.epoch & 1) * dst.parity_stride : 0LL);
    const unsigned n_rows = (unsigned)v.X * (unsigned)v.Y;
    const unsigned row0 = (row_block * blockDim.x + threadIdx.x) * EMIT_ROWS;
    const int lane = threadIdx.x & 31;
    const int k = z_lo >> 6, b0 = z_lo & 63, nb = z_hi - z_lo;
    const ull range = (nb >= 64 ? ~0ull : ((1ull << nb) - 1ull)) << b0;
    ull w[EMIT_ROWS];
    if (row0 + EMIT_ROWS <= n_rows) {
        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(f.face + row0));
        const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(f.face + row0) + 1);
        w[0] = ((ull)q0.y << 32) | q0.x; w[1] = ((ull)q0.w << 32) | q0.z;
        w[2] = ((ull)q1.y << 32) | q1.x; w[3] = ((ull)q1.w << 32) | q1.z;
    } else {
#pragma unroll
        for (int r = 0; r < EMIT_ROWS; ++r) w[r] = row0 + r < n_rows ? f.face[row0 + r] : 0ull;
    }
    ull any = 0ull;
#pragma unroll
    for (int r = 0; r < EMIT_ROWS; ++r) { w[r] &= range; any |= w[r]; }
    if (!__any_sync(0xffffffffu, any != 0ull)) return false;
    // one atomicAdd per warp (a per-run atomic on the single counter would serialise in L2)
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < EMIT_ROWS; ++r) cnt += __popcll(w[r] & ~(w[r] << 1));
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31) base = atomicAdd(runs, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    // The root of a run is a 3-load pointer chase in global memory.  Runs are first queued in shared memory
    // (start voxel, length), then taken one per lane, so a warp's chases overlap instead of following the
    // row-by-row order in which they were found (a warp of 128 rows holds ~1.5 runs, in different rows).
    int2* queue = s_runs[threadIdx.x >> 5];
    const bool queued = total <= EMIT_QUEUE;
    int at = incl - cnt;
#pragma unroll
    for (int r = 0; r < EMIT_ROWS; ++r) {
        const ull wr = w[r];
        const int gbase = (int)((row0 + r) * (unsigned)v.Z) + 64 * k;
        for (ull s = wr & ~(wr << 1); s; s &= s - 1, ++at) {
            const int p = __ffsll((long long)s) - 1;
            const ull tt = ~(wr >> p);
            const int len = tt ? __ffsll((long long)tt) - 1 : 64 - p;
            if (queued) {
                queue[at] = make_int2(gbase + p, len);
            } else {  // dense face: emit in place
                const int slot = base + at, root = gfind(v.parent, gbase + p);
                if (slot < cap) {
                    tri[3 * slot] = gbase + p;
                    tri[3 * slot + 1] = len;
                    tri[3 * slot + 2] = root;
                } else {
                    atomicOr(status, SKB_STATUS_ROOT_OVERFLOW);
                }
            }
        }
    }
    if (!queued) return cnt > 0;
    __syncwarp();
    for (int i = lane; i < total; i += 32) {
        const int2 e = queue[i];
        const int slot = base + i, root = gfind(v.parent, e.x);
        if (slot < cap) {
            tri[3 * slot] = e.x;
            tri[3 * slot + 1] = e.y;
            tri[3 * slot + 2] = root;
        } else {
            atomicOr(status, SKB_STATUS_ROOT_OVERFLOW);
        }
    }
    return lane < total;
}

__global__ void __launch_bounds__(256) shard_emit_runs_kernel(CclView v, EmitFace f0, EmitFace f1, int cap, unsigned* status) {
    __shared__ int2 s_runs[8][EMIT_QUEUE];
    const EmitFace& f = blockIdx.y ? f1 : f0;
    // Signalling from the kernel's last CTA (f.done != NULL) was measured and NOT adopted: counting 4096 CTAs per face on
    // one address serialises in L2 (30 -> 74 us), and a bounded grid that walks the row blocks loses as much to its
    // serial iterations (58 us for two faces).  The one-warp signal kernel behind this one costs 6 us.
    const unsigned n_row_blocks = ((unsigned)v.X * (unsigned)v.Y + blockDim.x * EMIT_ROWS - 1) / (blockDim.x * EMIT_ROWS);
    bool stored = false;
    for (unsigned rb = blockIdx.x; rb < n_row_blocks; rb += gridDim.x) {
        stored |= emit_face_rows(v, f, cap, status, s_runs, rb);
        __syncwarp();  // the warp's queue is reused by its next row block
    }
    if (f.done == nullptr) return;
    // my triples (stored into the neighbour's memory) become visible system-wide before my CTA counts itself as finished
    if (stored) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int finished = atomicAdd(f.done, 1);
        if (finished == (int)gridDim.x - 1) {  // every CTA of this face has counted itself: the run count is final
            __threadfence();
            *f.done = 0;
            const int e = *f.dst.epoch;
            f.remote_runs[(long long)(e & 1) * f.dst.parity_stride] = *reinterpret_cast<volatile int*>(f.dst.counter);
            __threadfence_system();
            st_release_sys(f.remote_flag, e);
        }
    }
}

// peer transport: publishes my face's run count in the neighbour's buffer, then releases its flag.  The
// triples were stored by the previous kernel on this stream, so they are ordered before the flag.
struct SignalRuns {
    const int* local_cnt;  // NULL: no such face
    int* remote_runs;
    int* remote_flag;
};
__global__ void shard_signal_runs_kernel(SignalRuns s0, SignalRuns s1, long long parity_stride, const int* epoch) {
    if (threadIdx.x < 2) {
        const SignalRuns& s = threadIdx.x ? s1 : s0;
        if (s.local_cnt) {
            const int e = *epoch;
            s.remote_runs[(long long)(e & 1) * parity_stride] = *s.local_cnt;
            __threadfence_system();
            st_release_sys(s.remote_flag, e);
        }
    }
}

// flag != NULL (peer transport): every CTA first waits until the neighbour's runs of this pass have landed
// where the (my root, neighbour root) pairs found while ingesting the UPPER neighbour's runs go: a run that starts in
// the neighbour's first plane z1 touches my last plane iff my voxel below it is foreground
struct PairSink {
    int* exch;  // NULL: no pairing (lower neighbour's runs) — [n_roots, n_pairs, roots[cap_roots], pairs[2*cap_pairs]]
    int cap_roots, cap_pairs;
    int z1;     // first plane of the upper neighbour
    int nk;     // words per row of my compact bit mask
};

__global__ void __launch_bounds__(256) shard_ingest_runs_kernel(CclView v, const int* runs, int cap, ull* __restrict__ halo,
                                                               const int* flag, const int* epoch, long long parity_stride,
                                                               PairSink ps) {
    if (flag) {
        const int e = *epoch;
        if (threadIdx.x == 0) spin_until(flag, e, v.status);
        __syncthreads();
        runs += (long long)(e & 1) * parity_stride;
    }
    const int n = min(runs[0], cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int s = runs[3 + 3 * i], len = runs[4 + 3 * i], root = runs[5 + 3 * i];
        const unsigned rowi = (unsigned)s / (unsigned)v.Z;
        const int z = (int)((unsigned)s - rowi * (unsigned)v.Z);
        const ull m = (len >= 64 ? ~0ull : ((1ull << len) - 1ull)) << (z & 63);
        atomicOr(&halo[rowi], m);
        for (int j = 0; j < len; ++j) v.parent[s + j] = root;
        v.parent[root] = root;  // the neighbour's root becomes a node of my union-find (idempotent)
        if (ps.exch && z == ps.z1 && (v.bits[(size_t)rowi * ps.nk + (ps.nk - 1)] >> 63)) {
            const int a = gfind(v.parent, s - 1);  // my voxel right below the run's first voxel
            const int slot = atomicAdd(ps.exch + 1, 1);
            if (slot < ps.cap_pairs) {
                ps.exch[2 + ps.cap_roots + 2 * slot] = a;
                ps.exch[3 + ps.cap_roots + 2 * slot] = root;
            } else {
                atomicOr(v.status, SKB_STATUS_ROOT_OVERFLOW);
            }
        }
    }
}

// exch = [n_roots, n_pairs, roots[cap_roots], pairs[2*cap_pairs]]
__global__ void __launch_bounds__(256) shard_pack_roots_kernel(CclView v, int* __restrict__ exch, int cap_roots,
                                                              unsigned* status) {
    const unsigned n = v.hdr->n_global_roots;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        exch[0] = (int)min(n, (unsigned)cap_roots);
        if (n > (unsigned)cap_roots) atomicOr(status, SKB_STATUS_ROOT_OVERFLOW);
    }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n && i < (unsigned)cap_roots; i += gridDim.x * blockDim.x)
        exch[2 + i] = v.groots[i];
}

__global__ void __launch_bounds__(256) shard_boundary_pairs_kernel(CclView v, const ull* __restrict__ halo_hi,
                                                                  int* __restrict__ exch, int cap_roots, int cap_pairs,
                                                                  unsigned* status) {
    // EMIT_ROWS consecutive rows per thread, two 16-byte loads of the halo words: ~99 % of the rows have no
    // foreground in the neighbour's first plane and leave after the loads
    const unsigned n_rows = (unsigned)v.X * (unsigned)v.Y;
    const unsigned row0 = (blockIdx.x * blockDim.x + threadIdx.x) * EMIT_ROWS;
    if (row0 >= n_rows) return;
    ull h[EMIT_ROWS];
    if (row0 + EMIT_ROWS <= n_rows) {
        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(halo_hi + row0));
        const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(halo_hi + row0) + 1);
        h[0] = ((ull)q0.y << 32) | q0.x; h[1] = ((ull)q0.w << 32) | q0.z;
        h[2] = ((ull)q1.y << 32) | q1.x; h[3] = ((ull)q1.w << 32) | q1.z;
    } else {
#pragma unroll
        for (int r = 0; r < EMIT_ROWS; ++r) h[r] = row0 + r < n_rows ? halo_hi[row0 + r] : 0ull;
    }
    const int z1 = v.z_off + v.Zl;  // first plane of the upper neighbour
#pragma unroll
    for (int r = 0; r < EMIT_ROWS; ++r) {
        if (!(h[r] & 1ull)) continue;
        const unsigned rowi = row0 + r;
        if (!(v.bits[(size_t)rowi * v.nk + (v.nk - 1)] >> 63)) continue;
        const int mine = (int)(rowi * (unsigned)v.Z) + z1 - 1;
        const int a = gfind(v.parent, mine), b = v.parent[mine + 1];
        const int slot = atomicAdd(exch + 1, 1);
        if (slot < cap_pairs) {
            exch[2 + cap_roots + 2 * slot] = a;
            exch[3 + cap_roots + 2 * slot] = b;
        } else {
            atomicOr(status, SKB_STATUS_ROOT_OVERFLOW);
        }
    }
}

struct MergeView {
    const int* gathered;  // world x stride ints (peer transport: copy 0; copy 1 is parity_stride further)
    int world, rank, stride, cap_roots, cap_pairs;
    const int* epoch;     // peer transport only, else NULL
    const int* flags;     // peer transport only: world flag words, SKB_FLAG_STRIDE apart
    long long parity_stride;
};

__device__ __forceinline__ const int* merge_base(const MergeView& m) {
    return m.gathered + (m.epoch ? (long long)(*m.epoch & 1) * m.parity_stride : 0LL);
}

// The merge kernels run on a grid (MERGE_BLOCKS, world): row y of the grid walks rank y's root list or pair
// list — only the entries that exist, not the capacity (8 ranks x 2^18 slots for ~4 K roots each).
constexpr int MERGE_BLOCKS = 48;
struct MergeList {
    const int* items;
    int n;
};
__device__ __forceinline__ MergeList merge_list(const MergeView& m, const int* base, bool pairs) {
    const int* e = base + (size_t)blockIdx.y * m.stride;
    MergeList l;
    l.n = min(pairs ? e[1] : e[0], pairs ? m.cap_pairs : m.cap_roots);
    l.items = e + 2 + (pairs ? m.cap_roots : 0);
    return l;
}

__global__ void __launch_bounds__(256) shard_merge_init_kernel(CclView v, MergeView m, int label_base) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {  // the mark kernel re-lists the roots; the rank kernel reads label_base
        v.hdr->n_global_roots = 0u;
        v.hdr->label_base = label_base;
    }
    if (m.flags) {  // peer transport: the first merge kernel waits for every rank's payload of this pass
        if ((int)threadIdx.x < m.world) spin_until(m.flags + threadIdx.x * SKB_FLAG_STRIDE, *m.epoch, v.status);
        __syncthreads();
    }
    if ((int)blockIdx.y == m.rank) return;
    const MergeList l = merge_list(m, merge_base(m), false);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < l.n; i += gridDim.x * blockDim.x)
        v.parent[l.items[i]] = l.items[i];  // foreign roots join my union-find as singletons
}

__global__ void __launch_bounds__(256) shard_merge_union_kernel(CclView v, MergeView m) {
    const MergeList l = merge_list(m, merge_base(m), true);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < l.n; i += gridDim.x * blockDim.x)
        gunion(v.parent, l.items[2 * i], l.items[2 * i + 1]);
}

// global roots among all ranks' roots -> bitmap + chunk histogram + list (reuses groots: the local list
// has already been shipped in the exchange buffer)
__global__ void __launch_bounds__(256) shard_merge_mark_kernel(CclView v, MergeView m) {
    const MergeList l = merge_list(m, merge_base(m), false);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < l.n; i += gridDim.x * blockDim.x) {
        const int root = l.items[i];
        if (gfind(v.parent, root) != root) continue;
        unsigned slot = atomicAdd(&v.hdr->n_global_roots, 1u);
        if (slot >= (unsigned)v.capacity) {  // not listed -> could not be cleared afterwards: do not mark it
            atomicOr(v.status, SKB_STATUS_ROOT_OVERFLOW);
            continue;
        }
        v.groots[slot] = root;
        int bit;
        long long wi = word_of_voxel(v, root, &bit);
        atomicOr(&v.rootbits[wi], 1ull << bit);
        atomicAdd(&v.chunks[wi >> 6], 1);
    }
}

// every listed root takes the label code of its global root (codes are negative, indices are not); the
// root bitmap words the global roots touched are zeroed for the next pass (ccl_clear_rootbits_kernel's job)
__global__ void __launch_bounds__(256) shard_publish_roots_kernel(CclView v, MergeView m) {
    if (blockIdx.y == 0) {
        const unsigned n_g = min(v.hdr->n_global_roots, (unsigned)v.capacity);
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_g; i += gridDim.x * blockDim.x) {
            int bit;
            v.rootbits[word_of_voxel(v, v.groots[i], &bit)] = 0ull;
        }
    }
    const MergeList l = merge_list(m, merge_base(m), false);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < l.n; i += gridDim.x * blockDim.x) {
        const int root = l.items[i];
        int a = root, p = gload(v.parent + a);
        while (p >= 0 && p != a) { a = p; p = gload(v.parent + a); }
        if (p < 0 && a != root) v.parent[root] = p;
    }
}

// ---- the merge as ONE cooperative kernel --------------------------------------------------------------------
// Round 1 ran the merge as eight dependent launches (init, union, mark, scan x2, rank, publish x2): 56 us of an 8-GPU
// pass for a few thousand roots and pairs — almost all of it launch-to-launch latency.  Here the same phases run
// inside one cooperatively launched grid (one 1024-thread CTA per SM, every CTA resident), separated by grid-wide
// barriers (~1.5 us each).  The phases are the kernels above, verbatim, with the (block, rank-row) indices of their
// grids mapped onto this grid's CTAs.
constexpr int MERGE_FUSED_THREADS = 1024;
constexpr int MERGE_FUSED_CTAS = 32;

template <typename F>
__device__ __forceinline__ void merge_for_each(const MergeView& m, const int* base, bool pairs, bool skip_own, F&& f) {
    const int per = gridDim.x / m.world > 0 ? gridDim.x / m.world : 1;  // CTAs per rank-row
    for (int row = blockIdx.x / per; row < m.world; row += (gridDim.x + per - 1) / per) {
        if (skip_own && row == m.rank) continue;
        const int* e = base + (size_t)row * m.stride;
        const int n = min(pairs ? e[1] : e[0], pairs ? m.cap_pairs : m.cap_roots);
        const int* items = e + 2 + (pairs ? m.cap_roots : 0);
        for (int i = (blockIdx.x % per) * blockDim.x + threadIdx.x; i < n; i += per * blockDim.x) f(items, i);
    }
}

__global__ void __launch_bounds__(MERGE_FUSED_THREADS, 1) shard_merge_fused_kernel(CclView v, MergeView m, int label_base) {
    cg::grid_group grid = cg::this_grid();
    __shared__ int warp_sums[32];
    __shared__ int total;
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    // phase 0: wait for every rank's payload of this pass (peer transport), reset the header fields the phases append to
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            v.hdr->n_global_roots = 0u;
            v.hdr->label_base = label_base;
        }
        if (m.flags && (int)threadIdx.x < m.world) spin_until(m.flags + threadIdx.x * SKB_FLAG_STRIDE, *m.epoch, v.status);
        __threadfence();
    }
    grid.sync();
    const int* base = merge_base(m);
    // phase 1: foreign roots join my union-find as singletons
    merge_for_each(m, base, false, true, [&](const int* items, int i) { v.parent[items[i]] = items[i]; });
    grid.sync();
    // phase 2: every rank's face pairs
    merge_for_each(m, base, true, false, [&](const int* items, int i) { gunion(v.parent, items[2 * i], items[2 * i + 1]); });
    grid.sync();
    // phase 3: global roots among all ranks' roots -> list + bitmap + chunk histogram
    merge_for_each(m, base, false, false, [&](const int* items, int i) {
        const int root = items[i];
        if (gfind(v.parent, root) != root) return;
        unsigned slot = atomicAdd(&v.hdr->n_global_roots, 1u);
        if (slot >= (unsigned)v.capacity) {
            atomicOr(v.status, SKB_STATUS_ROOT_OVERFLOW);
            return;
        }
        v.groots[slot] = root;
        int bit;
        long long wi = word_of_voxel(v, root, &bit);
        atomicOr(&v.rootbits[wi], 1ull << bit);
        atomicAdd(&v.chunks[wi >> 6], 1);
    });
    grid.sync();
    // phase 4 / 5: two-level exclusive scan of the chunk histogram
    for (long long t = blockIdx.x; t < v.n_scan_tiles; t += gridDim.x) {
        ccl_scan_tile_body(v, t, warp_sums, &total);
        __syncthreads();
    }
    grid.sync();
    if (blockIdx.x == 0) ccl_scan_top_body(v, warp_sums, &total);
    grid.sync();
    // phase 6: raster rank of every global root -> label code
    const unsigned n_g = min(v.hdr->n_global_roots, (unsigned)v.capacity);
    for (unsigned i = tid; i < n_g; i += nthr) ccl_rank_root(v, v.groots[i]);
    grid.sync();
    // phase 7: every listed root takes the code of its global root; the root bitmap is left zeroed for the next pass
    for (unsigned i = tid; i < n_g; i += nthr) {
        int bit;
        v.rootbits[word_of_voxel(v, v.groots[i], &bit)] = 0ull;
    }
    merge_for_each(m, base, false, false, [&](const int* items, int i) {
        const int root = items[i];
        int a = root, p = gload(v.parent + a);
        while (p >= 0 && p != a) { a = p; p = gload(v.parent + a); }
        if (p < 0 && a != root) v.parent[root] = p;
    });
    grid.sync();
    // phase 8: tile roots that are not slab roots take the code of their root
    const unsigned n_t = min(v.hdr->n_tile_roots, (unsigned)v.capacity);
    for (unsigned i = tid; i < n_t; i += nthr) {
        const int r = v.tile_roots[i], g = v.flat[i];
        if (g != r) v.parent[r] = v.parent[g];
    }
}

static int shard_common(const char* who, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl) {
    int rc = skb_check_volume(X, Y, Z, who);
    if (rc) return rc;
    if (Z % 64 != 0 || z_off % 64 != 0 || Zl % 64 != 0 || Zl <= 0 || z_off < 0 || z_off + Zl > Z) {
        skb_set_error("%s: Z, z_off and Zl must be multiples of 64 with the slab inside the volume", who);
        return SKB_E_ARG;
    }
    return SKB_OK;
}

static CclView slab_view(const SkbCclLayout& L, void* ws, int64_t capacity, int64_t z_off, int64_t Zl, uint32_t* status) {
    CclView v = skb_ccl_make_view(L, ws, 0, capacity, status, nullptr);
    v.z_off = (int)z_off; v.Zl = (int)Zl;
    v.k0 = (int)(z_off / 64); v.nk = (int)(Zl / 64);
    v.nk_shift = shift_of(v.nk);
    v.n_words = (long long)L.X * L.Y * v.nk;  // of the slab's compact bit mask
    // compact copies of every row's first / last word; a one-word-deep slab needs none: both ARE the bit mask
    v.face_lo = v.nk == 1 ? nullptr : reinterpret_cast<ull*>(static_cast<char*>(ws) + L.off_face_lo);
    v.face_hi = v.nk == 1 ? nullptr : reinterpret_cast<ull*>(static_cast<char*>(ws) + L.off_face_hi);
    return v;
}

extern "C" int skb_shard_label_local(const void* mask, int mask_dtype, int64_t X, int64_t Y, int64_t Z, int64_t z_off,
                                     int64_t Zl, int64_t capacity, void* workspace, size_t workspace_bytes,
                                     uint32_t* status, int flags, void* stream) {
    int rc = shard_common("skb_shard_label_local", X, Y, Z, z_off, Zl);
    if (rc) return rc;
    SKB_REQUIRE(mask && workspace && status, "skb_shard_label_local: NULL pointer");
    SKB_REQUIRE(mask_dtype == SKB_U8 || mask_dtype == SKB_I16, "skb_shard_label_local: mask dtype must be u8 or i16");
    SKB_REQUIRE(capacity > 0 && capacity <= 0x7fffffff, "skb_shard_label_local: bad capacity");
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, capacity);
    if (workspace_bytes < L.total) {
        skb_set_error("skb_shard_label_local: workspace %zu < required %zu bytes", workspace_bytes, L.total);
        return SKB_E_WORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CclView v = slab_view(L, workspace, capacity, z_off, Zl, status);
    SkbCclHeader h = {};
    h.capacity = (int)capacity;
    h.dims[0] = L.X; h.dims[1] = L.Y; h.dims[2] = L.Z;
    skb_ccl_launch_pack_and_tile(mask, mask_dtype, v, L, h, flags, st);
    SKB_LAUNCH_CHECK("skb_shard_label_local (pack/tile)");
    if ((flags & SKB_CCL_PHASE_PACK) && !(flags & SKB_CCL_PHASE_LABEL)) return SKB_OK;
    skb_ccl_launch_boundary(v, L, 8, 8, st);
    shard_local_roots_kernel<<<148 * 4, 256, 0, st>>>(v);
    SKB_LAUNCH_CHECK("skb_shard_label_local");
    return SKB_OK;
}

extern "C" int skb_shard_emit_runs(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                                   int face_is_high, int64_t halo, int32_t* runs, int64_t cap, uint32_t* status,
                                   void* stream) {
    int rc = shard_common("skb_shard_emit_runs", X, Y, Z, z_off, Zl);
    if (rc) return rc;
    SKB_REQUIRE(workspace && runs && status && cap > 0, "skb_shard_emit_runs: bad argument");
    SKB_REQUIRE(halo >= 1 && halo <= 64 && halo <= Zl, "skb_shard_emit_runs: halo must be 1..64 planes and fit the slab");
    SKB_REQUIRE(face_is_high == 0 || face_is_high == 1, "skb_shard_emit_runs: face_is_high must be 0 or 1");
    const int64_t z_lo = face_is_high ? z_off + Zl - halo : z_off, z_hi = face_is_high ? z_off + Zl : z_off + halo;
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, 1);
    CclView v = skb_ccl_make_view(L, workspace, 0, 1, status, nullptr);
    const ull* face = face_words(L, workspace, face_is_high, Zl);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(runs, 0, 3 * sizeof(int32_t), st);
    unsigned nb = (unsigned)(((long long)X * Y + 256 * EMIT_ROWS - 1) / (256 * EMIT_ROWS));
    EmitFace f = {face, (int)z_lo, (int)z_hi, {runs, runs + 3, nullptr, 0}};
    shard_emit_runs_kernel<<<nb, 256, 0, st>>>(v, f, f, (int)cap, status);
    SKB_LAUNCH_CHECK("shard_emit_runs_kernel");
    return SKB_OK;
}

// Zeroes exactly the halo words the PREVIOUS pass's ingest set, by walking that pass's run list (still intact in
// the receive buffer) — instead of a 33 MB memset of all X*Y words per face and pass.
__global__ void __launch_bounds__(256) shard_clear_halo_kernel(const int* runs, int cap, ull* __restrict__ halo, unsigned Z,
                                                              const int* epoch, long long parity_stride) {
    if (epoch) runs += (long long)((*epoch + 1) & 1) * parity_stride;  // the copy the previous pass used
    const int n = min(runs[0], cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        halo[(unsigned)runs[3 + 3 * i] / Z] = 0ull;
}

extern "C" int skb_shard_clear_halo(int64_t Z, const int32_t* prev_runs, int64_t cap, uint64_t* halo_words, void* stream) {
    SKB_REQUIRE(Z > 0 && prev_runs && halo_words && cap > 0, "skb_shard_clear_halo: bad argument");
    shard_clear_halo_kernel<<<148, 256, 0, static_cast<cudaStream_t>(stream)>>>(prev_runs, (int)cap, reinterpret_cast<ull*>(halo_words),
                                                                                (unsigned)Z, nullptr, 0);
    SKB_LAUNCH_CHECK("shard_clear_halo_kernel");
    return SKB_OK;
}

static PairSink pair_sink(int32_t* exchange, int64_t cap_roots, int64_t cap_pairs, int64_t z_off, int64_t Zl) {
    PairSink ps = {exchange, (int)cap_roots, (int)cap_pairs, (int)(z_off + Zl), (int)(Zl / 64)};
    return ps;
}

extern "C" int skb_shard_ingest_runs(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                                     const int32_t* runs, int64_t cap, uint64_t* halo_words_zeroed, int32_t* exchange,
                                     int64_t cap_roots, int64_t cap_pairs, uint32_t* status, void* stream) {
    int rc = shard_common("skb_shard_ingest_runs", X, Y, Z, z_off, Zl);
    if (rc) return rc;
    SKB_REQUIRE(workspace && runs && halo_words_zeroed && status && cap > 0, "skb_shard_ingest_runs: bad argument");
    SKB_REQUIRE(!exchange || (cap_roots > 0 && cap_pairs > 0), "skb_shard_ingest_runs: bad exchange capacities");
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, 1);
    CclView v = skb_ccl_make_view(L, workspace, 0, 1, status, nullptr);
    shard_ingest_runs_kernel<<<148 * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        v, runs, (int)cap, reinterpret_cast<ull*>(halo_words_zeroed), nullptr, nullptr, 0,
        pair_sink(exchange, cap_roots, cap_pairs, z_off, Zl));
    SKB_LAUNCH_CHECK("shard_ingest_runs_kernel");
    return SKB_OK;
}

extern "C" int skb_shard_boundary_pairs(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                                        int64_t capacity, const uint64_t* halo_hi, int32_t* exchange, int64_t cap_roots,
                                        int64_t cap_pairs, uint32_t* status, void* stream) {
    int rc = shard_common("skb_shard_boundary_pairs", X, Y, Z, z_off, Zl);
    if (rc) return rc;
    SKB_REQUIRE(workspace && exchange && status && cap_roots > 0 && cap_pairs > 0, "skb_shard_boundary_pairs: bad argument");
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, capacity);
    CclView v = slab_view(L, workspace, capacity, z_off, Zl, status);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(exchange, 0, 2 * sizeof(int32_t), st);
    shard_pack_roots_kernel<<<148, 256, 0, st>>>(v, exchange, (int)cap_roots, status);
    if (halo_hi && z_off + Zl < Z) {
        unsigned nb = (unsigned)(((long long)X * Y + 256 * EMIT_ROWS - 1) / (256 * EMIT_ROWS));
        shard_boundary_pairs_kernel<<<nb, 256, 0, st>>>(v, reinterpret_cast<const ull*>(halo_hi), exchange, (int)cap_roots,
                                                       (int)cap_pairs, status);
    }
    SKB_LAUNCH_CHECK("skb_shard_boundary_pairs");
    return SKB_OK;
}

static int launch_merge(const CclView& v, const SkbCclLayout& L, const MergeView& m, int32_t label_base, cudaStream_t st);

// word k0 (low face) / k0+nk-1 (high face) of every row of a slab `Zl` planes deep, contiguous
static const ull* face_words(const SkbCclLayout& L, const void* workspace, int high, int64_t Zl) {
    const char* base = static_cast<const char*>(workspace);
    if (Zl == 64) return reinterpret_cast<const ull*>(base + L.off_bits);  // the compact bit mask itself
    return reinterpret_cast<const ull*>(base + (high ? L.off_face_hi : L.off_face_lo));
}

extern "C" int skb_shard_merge(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t capacity, const int32_t* gathered,
                               int world, int rank, int64_t cap_roots, int64_t cap_pairs, int32_t label_base,
                               int32_t* ncomp, uint32_t* status, void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_shard_merge");
    if (rc) return rc;
    SKB_REQUIRE(workspace && gathered && status && world >= 1 && rank >= 0 && rank < world && label_base >= 0,
                "skb_shard_merge: bad argument");
    SKB_REQUIRE((long long)world * cap_roots < (1LL << 31) && (long long)world * cap_pairs < (1LL << 31), "skb_shard_merge: capacities too large");
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, capacity);
    CclView v = skb_ccl_make_view(L, workspace, 0, capacity, status, ncomp);
    MergeView m = {};
    m.gathered = gathered; m.world = world; m.rank = rank;
    m.cap_roots = (int)cap_roots; m.cap_pairs = (int)cap_pairs;
    m.stride = 2 + (int)cap_roots + 2 * (int)cap_pairs;
    return launch_merge(v, L, m, label_base, static_cast<cudaStream_t>(stream));
}

static int launch_merge(const CclView& v, const SkbCclLayout& L, const MergeView& m, int32_t label_base, cudaStream_t st) {
    // The one-kernel form is opt-in (SKB_SHARD_FUSED=1).  Measured on a 1/8 slab of the headline volume (ncu launch list,
    // profiles/r02_launches_shard8_emulated.txt): 64 us with one CTA per SM, 73 us with 32 CTAs, against 56 us for the sum
    // of the eight launches it replaces — eight grid-wide barriers of 1024-thread CTAs cost more than eight launch gaps.
    static int fused_ok = -1;
    if (fused_ok < 0) {
        const char* e = getenv("SKB_SHARD_FUSED");
        int dev = 0, coop = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shard_merge_fused_kernel, MERGE_FUSED_THREADS, 0);
        fused_ok = ((e && e[0] == '1') && coop && per_sm >= 1) ? 1 : 0;
    }
    if (fused_ok) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        CclView vv = v;
        MergeView mm = m;
        int lb = label_base;
        void* args[] = {&vv, &mm, &lb};
        // a few thousand roots and pairs: 32 CTAs have more than enough threads, and a grid barrier among 32 CTAs costs a
        // fraction of one among 148 (measured with all SMs: 64 us for the kernel, more than the eight launches it replaced)
        const int ctas = sms < MERGE_FUSED_CTAS ? sms : MERGE_FUSED_CTAS;
        cudaError_t rc = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(shard_merge_fused_kernel), dim3((unsigned)ctas),
                                                     dim3(MERGE_FUSED_THREADS), args, 0, st);
        if (rc == cudaSuccess) return SKB_OK;
        (void)cudaGetLastError();
        fused_ok = 0;  // not launchable here: the eight-launch form from now on
    }
    const dim3 per_rank(MERGE_BLOCKS, m.world);
    shard_merge_init_kernel<<<per_rank, 256, 0, st>>>(v, m, label_base);
    shard_merge_union_kernel<<<per_rank, 256, 0, st>>>(v, m);
    shard_merge_mark_kernel<<<per_rank, 256, 0, st>>>(v, m);
    skb_ccl_launch_scan_and_rank(v, L, st);
    shard_publish_roots_kernel<<<per_rank, 256, 0, st>>>(v, m);  // also clears the root bitmap
    skb_ccl_launch_publish(v, st);
    SKB_LAUNCH_CHECK("skb_shard_merge");
    return SKB_OK;
}

// ==========================================================================================
// Peer transport: the two exchanges are plain stores into the consumer GPU's mailbox over NVLink
// plus release/acquire flags (include/skoots_b200.h (e')).
// ==========================================================================================
struct Mailbox {
    SkbMailboxLayout M;
    char* base;
    int* ints(size_t off) const { return reinterpret_cast<int*>(base + off); }
    int* epoch() const { return ints(M.off_epoch); }
    int* cnt(int hi) const { return ints(M.off_cnt) + hi; }
    int* flag(int hi) const { return ints(hi ? M.off_flag_hi : M.off_flag_lo); }
    int* flag_gather(int r) const { return ints(M.off_flag_gather) + (size_t)r * SKB_FLAG_STRIDE; }
    int* recv(int hi) const { return ints(hi ? M.off_recv_hi : M.off_recv_lo); }
    int* gathered() const { return ints(M.off_gathered); }
};

static int mailbox_args(const char* who, int world, int64_t cap_runs, int64_t cap_roots, int64_t cap_pairs) {
    if (world < 1 || world > SKB_MAX_WORLD || cap_runs <= 0 || cap_roots <= 0 || cap_pairs <= 0 ||
        cap_runs >= (1LL << 28) || (long long)world * (2 + cap_roots + 2 * cap_pairs) >= (1LL << 31)) {
        skb_set_error("%s: bad mailbox geometry (world 1..%d, capacities > 0 and < 2^31 ints in total)", who, SKB_MAX_WORLD);
        return SKB_E_ARG;
    }
    return SKB_OK;
}

static Mailbox mailbox_at(void* base, int world, int64_t cap_runs, int64_t cap_roots, int64_t cap_pairs) {
    Mailbox b;
    b.M = skb_mailbox_layout(world, cap_runs, cap_roots, cap_pairs);
    b.base = static_cast<char*>(base);
    return b;
}

extern "C" size_t skb_shard_mailbox_bytes(int world, int64_t cap_runs, int64_t cap_roots, int64_t cap_pairs) {
    if (mailbox_args("skb_shard_mailbox_bytes", world, cap_runs, cap_roots, cap_pairs)) return 0;
    return skb_mailbox_layout(world, cap_runs, cap_roots, cap_pairs).total;
}

__global__ void shard_begin_kernel(int* epoch, int* cnt, int* exch) {
    if (threadIdx.x == 0) {
        *epoch += 1;
        cnt[0] = 0;
        cnt[1] = 0;
        if (exch) { exch[0] = 0; exch[1] = 0; }  // [n_roots (unused by the peer transport), n_pairs]
    }
}

// both faces' halo words of the previous pass in one launch (blockIdx.y = face)
__global__ void __launch_bounds__(256) shard_clear_halos_kernel(const int* runs_lo, const int* runs_hi, int cap, ull* __restrict__ halo_lo,
                                                               ull* __restrict__ halo_hi, unsigned Z, const int* epoch,
                                                               long long parity_stride) {
    const int* runs = blockIdx.y ? runs_hi : runs_lo;
    ull* halo = blockIdx.y ? halo_hi : halo_lo;
    if (!halo) return;
    runs += (long long)((*epoch + 1) & 1) * parity_stride;  // the copy the previous pass used
    const int n = min(runs[0], cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        halo[(unsigned)runs[3 + 3 * i] / Z] = 0ull;
}

extern "C" int skb_shard_begin(void* mailbox, int world, int64_t cap_runs, int64_t cap_roots, int64_t cap_pairs,
                               void* stream) {
    return skb_shard_begin_pass(mailbox, world, cap_runs, cap_roots, cap_pairs, 1, nullptr, nullptr, nullptr, stream);
}

extern "C" int skb_shard_begin_pass(void* mailbox, int world, int64_t cap_runs, int64_t cap_roots, int64_t cap_pairs, int64_t Z,
                                    uint64_t* halo_lo, uint64_t* halo_hi, int32_t* exchange, void* stream) {
    int rc = mailbox_args("skb_shard_begin", world, cap_runs, cap_roots, cap_pairs);
    if (rc) return rc;
    SKB_REQUIRE(mailbox && Z > 0, "skb_shard_begin: NULL mailbox");
    Mailbox mb = mailbox_at(mailbox, world, cap_runs, cap_roots, cap_pairs);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    shard_begin_kernel<<<1, 32, 0, st>>>(mb.epoch(), mb.cnt(0), exchange);
    if (halo_lo || halo_hi)
        shard_clear_halos_kernel<<<dim3(74, 2), 256, 0, st>>>(mb.recv(0), mb.recv(1), (int)cap_runs, reinterpret_cast<ull*>(halo_lo),
                                                             reinterpret_cast<ull*>(halo_hi), (unsigned)Z, mb.epoch(), mb.M.runs_ints);
    SKB_LAUNCH_CHECK("shard_begin_kernel");
    return SKB_OK;
}

extern "C" int skb_shard_clear_halo_peer(int64_t Z, void* mailbox, int from_high, int world, int64_t cap_runs,
                                         int64_t cap_roots, int64_t cap_pairs, uint64_t* halo_words, void* stream) {
    int rc = mailbox_args("skb_shard_clear_halo_peer", world, cap_runs, cap_roots, cap_pairs);
    if (rc) return rc;
    SKB_REQUIRE(Z > 0 && mailbox && halo_words && (from_high == 0 || from_high == 1), "skb_shard_clear_halo_peer: bad argument");
    Mailbox me = mailbox_at(mailbox, world, cap_runs, cap_roots, cap_pairs);
    shard_clear_halo_kernel<<<148, 256, 0, static_cast<cudaStream_t>(stream)>>>(me.recv(from_high), (int)cap_runs,
                                                                                reinterpret_cast<ull*>(halo_words), (unsigned)Z,
                                                                                me.epoch(), me.M.runs_ints);
    SKB_LAUNCH_CHECK("shard_clear_halo_kernel (peer)");
    return SKB_OK;
}

extern "C" int skb_shard_emit_runs_peer(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                                        int64_t halo, void* mailbox, void* lo_neighbour_mailbox,
                                        void* hi_neighbour_mailbox, int world, int64_t cap_runs, int64_t cap_roots,
                                        int64_t cap_pairs, uint32_t* status, void* stream) {
    int rc = shard_common("skb_shard_emit_runs_peer", X, Y, Z, z_off, Zl);
    if (rc) return rc;
    rc = mailbox_args("skb_shard_emit_runs_peer", world, cap_runs, cap_roots, cap_pairs);
    if (rc) return rc;
    SKB_REQUIRE(workspace && mailbox && status, "skb_shard_emit_runs_peer: NULL pointer");
    SKB_REQUIRE(halo >= 1 && halo <= 64 && halo <= Zl, "skb_shard_emit_runs_peer: halo must be 1..64 planes and fit the slab");
    if (!lo_neighbour_mailbox && !hi_neighbour_mailbox) return SKB_OK;
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, 1);
    CclView v = skb_ccl_make_view(L, workspace, 0, 1, status, nullptr);
    Mailbox me = mailbox_at(mailbox, world, cap_runs, cap_roots, cap_pairs);
    // my LOW face lands in the lower neighbour's recv_hi, my HIGH face in the upper neighbour's recv_lo
    EmitFace f[2] = {};
    int n = 0;
    for (int hi = 0; hi < 2; ++hi) {
        void* nbp = hi ? hi_neighbour_mailbox : lo_neighbour_mailbox;
        if (!nbp) continue;
        Mailbox nb = mailbox_at(nbp, world, cap_runs, cap_roots, cap_pairs);
        int* remote = nb.recv(hi ? 0 : 1);
        f[n].face = face_words(L, workspace, hi, Zl);
        f[n].z_lo = (int)(hi ? z_off + Zl - halo : z_off);
        f[n].z_hi = (int)(hi ? z_off + Zl : z_off + halo);
        f[n].dst = {me.cnt(hi), remote + 3, me.epoch(), me.M.runs_ints};
        f[n].done = me.cnt(2 + hi);  // counters 2, 3 of the same line: CTAs of my low / high face that have finished
        f[n].remote_runs = remote;
        f[n].remote_flag = nb.flag(hi ? 0 : 1);
        ++n;
    }
    if (n == 1) f[1] = f[0];
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned nblk = (unsigned)(((long long)X * Y + 256 * EMIT_ROWS - 1) / (256 * EMIT_ROWS));
    SignalRuns sg[2] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    for (int i = 0; i < n; ++i) {
        sg[i] = {f[i].dst.counter, f[i].remote_runs, f[i].remote_flag};
        f[i].done = nullptr;  // see the kernel: signalling from its last CTA is slower than the signal kernel
    }
    shard_emit_runs_kernel<<<dim3(nblk, n), 256, 0, st>>>(v, f[0], f[1], (int)cap_runs, status);
    shard_signal_runs_kernel<<<1, 32, 0, st>>>(sg[0], sg[1], me.M.runs_ints, me.epoch());
    SKB_LAUNCH_CHECK("skb_shard_emit_runs_peer");
    return SKB_OK;
}

extern "C" int skb_shard_ingest_runs_peer(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                                          void* mailbox, int from_high, int world, int64_t cap_runs, int64_t cap_roots,
                                          int64_t cap_pairs, uint64_t* halo_words_zeroed, int32_t* exchange,
                                          uint32_t* status, void* stream) {
    int rc = shard_common("skb_shard_ingest_runs_peer", X, Y, Z, z_off, Zl);
    if (rc) return rc;
    rc = mailbox_args("skb_shard_ingest_runs_peer", world, cap_runs, cap_roots, cap_pairs);
    if (rc) return rc;
    SKB_REQUIRE(workspace && mailbox && halo_words_zeroed && status, "skb_shard_ingest_runs_peer: NULL pointer");
    SKB_REQUIRE(from_high == 0 || from_high == 1, "skb_shard_ingest_runs_peer: from_high must be 0 or 1");
    SKB_REQUIRE(!exchange || from_high == 1, "skb_shard_ingest_runs_peer: pairs are found in the upper neighbour's runs only");
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, 1);
    CclView v = skb_ccl_make_view(L, workspace, 0, 1, status, nullptr);
    Mailbox me = mailbox_at(mailbox, world, cap_runs, cap_roots, cap_pairs);
    shard_ingest_runs_kernel<<<148 * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        v, me.recv(from_high), (int)cap_runs, reinterpret_cast<ull*>(halo_words_zeroed), me.flag(from_high), me.epoch(),
        me.M.runs_ints, pair_sink(exchange, cap_roots, cap_pairs, z_off, Zl));
    SKB_LAUNCH_CHECK("shard_ingest_runs_kernel (peer)");
    return SKB_OK;
}

struct PeerTable {
    int* gathered[SKB_MAX_WORLD];  // copy 0 of rank p's gather buffer
    int* flag[SKB_MAX_WORLD];      // rank p's flag word for payloads coming from me
};

// grid (PUSH_BLOCKS, world): CTA (.,p) stores my payload — [n_roots, n_pairs, roots, pairs] — into slot `rank` of rank p:
// the roots straight from the slab's root list (round 1 first packed them into the exchange buffer with one more
// kernel), the pairs from where the ingest appended them.  The last CTA to finish releases every rank's flag (round 1:
// a separate one-warp kernel).
constexpr int PUSH_BLOCKS = 4;
__global__ void __launch_bounds__(256) shard_push_kernel(CclView v, const int* __restrict__ exch, PeerTable T, int world, int rank,
                                                        int cap_roots, int cap_pairs, long long stride, long long parity_stride,
                                                        const int* epoch, int* done) {
    const int p = blockIdx.y;
    const int e = *epoch;
    int* dst = T.gathered[p] + (long long)(e & 1) * parity_stride + (long long)rank * stride;
    const unsigned n_all = v.hdr->n_global_roots;
    const int n_roots = (int)min(n_all, (unsigned)cap_roots), n_pairs = min(exch[1], cap_pairs);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if (tid == 0) {
        dst[0] = n_roots;
        dst[1] = n_pairs;
        if (p == 0 && n_all > (unsigned)cap_roots) atomicOr(v.status, SKB_STATUS_ROOT_OVERFLOW);
    }
    for (int i = tid; i < n_roots; i += nthr) dst[2 + i] = v.groots[i];
    const int* ps = exch + 2 + cap_roots;
    int* pd = dst + 2 + cap_roots;
    for (int i = tid; i < 2 * n_pairs; i += nthr) pd[i] = ps[i];
    __syncthreads();  // the CTA's stores happen-before thread 0's fence (fences are cumulative): one system fence per CTA
    __shared__ int last;
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(done, 1) == (int)(gridDim.x * gridDim.y) - 1;
    }
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *done = 0;
        if ((int)threadIdx.x < world) {
            __threadfence_system();
            st_release_sys(T.flag[threadIdx.x], e);
        }
    }
}

extern "C" int skb_shard_push(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t capacity, const int32_t* exchange, void* mailbox,
                              const uint64_t* peer_mailboxes, int world, int rank, int64_t cap_runs, int64_t cap_roots,
                              int64_t cap_pairs, uint32_t* status, void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_shard_push");
    if (rc) return rc;
    rc = mailbox_args("skb_shard_push", world, cap_runs, cap_roots, cap_pairs);
    if (rc) return rc;
    SKB_REQUIRE(workspace && exchange && mailbox && peer_mailboxes && status && rank >= 0 && rank < world, "skb_shard_push: bad argument");
    Mailbox me = mailbox_at(mailbox, world, cap_runs, cap_roots, cap_pairs);
    PeerTable T = {};
    for (int p = 0; p < world; ++p) {
        SKB_REQUIRE(peer_mailboxes[p] != 0, "skb_shard_push: NULL peer mailbox");
        Mailbox pm = mailbox_at(reinterpret_cast<void*>(static_cast<uintptr_t>(peer_mailboxes[p])), world, cap_runs, cap_roots, cap_pairs);
        T.gathered[p] = pm.gathered();
        T.flag[p] = pm.flag_gather(rank);
    }
    SKB_REQUIRE(capacity > 0 && capacity <= 0x7fffffff, "skb_shard_push: bad capacity");
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, capacity);  // the root list sits behind the capacity-sized arrays
    CclView v = skb_ccl_make_view(L, workspace, 0, capacity, status, nullptr);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long parity_stride = (long long)world * me.M.stride;
    shard_push_kernel<<<dim3(PUSH_BLOCKS, world), 256, 0, st>>>(v, exchange, T, world, rank, (int)cap_roots, (int)cap_pairs, me.M.stride,
                                                               parity_stride, me.epoch(), me.cnt(4));
    SKB_LAUNCH_CHECK("skb_shard_push");
    return SKB_OK;
}

extern "C" int skb_shard_merge_peer(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t capacity, void* mailbox,
                                    int world, int rank, int64_t cap_runs, int64_t cap_roots, int64_t cap_pairs,
                                    int32_t label_base, int32_t* ncomp, uint32_t* status, void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_shard_merge_peer");
    if (rc) return rc;
    rc = mailbox_args("skb_shard_merge_peer", world, cap_runs, cap_roots, cap_pairs);
    if (rc) return rc;
    SKB_REQUIRE(workspace && mailbox && status && rank >= 0 && rank < world && label_base >= 0, "skb_shard_merge_peer: bad argument");
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, capacity);
    CclView v = skb_ccl_make_view(L, workspace, 0, capacity, status, ncomp);
    Mailbox me = mailbox_at(mailbox, world, cap_runs, cap_roots, cap_pairs);
    MergeView m = {};
    m.gathered = me.gathered(); m.world = world; m.rank = rank;
    m.cap_roots = (int)cap_roots; m.cap_pairs = (int)cap_pairs;
    m.stride = (int)me.M.stride;
    m.epoch = me.epoch(); m.flags = me.flag_gather(0);
    m.parity_stride = (long long)world * me.M.stride;
    return launch_merge(v, L, m, label_base, static_cast<cudaStream_t>(stream));
}
