// Row f2 of the scope table: what the reference does to an instance mask after assembly.
//
//   skb_renumber          fastremap.renumber(mask, in_place=True)            skoots/lib/eval.py:304
//                         labels -> 1..N in order of first appearance in the C-order scan, 0 stays 0
//   skb_unique_index      gt.unique() / pred.unique() (sorted, > 0)          skoots/validate/lib.py:201-205
//   skb_contingency       every `logical_and(_a, _b).sum()` / `_a.sum()` of  skoots/validate/lib.py:211-226, 253-273
//                         the reference's O(N·M) loop, as ONE pass: a contingency table
//   skb_iou_dice          iou = I/(A+B-I), dice = 2I/(A+B), fp32 division of int counts (the reference divides
//                         0-dim int64 tensors: true_divide -> float32)
//   skb_accuracies_from_iou   skoots/validate/lib.py:170-187
//
// All of it is table-driven integer work on label values: labels must be non-negative and smaller than the
// caller's `table_size` (instance labels are 0..N+2 on this path); anything else sets SKB_STATUS_LABEL_RANGE.
// Streaming passes read 8 voxels per thread with 16-byte loads and touch the tables only at RUN STARTS (a voxel
// whose label differs from its predecessor's): instance masks are piecewise constant along z, so the table
// atomics shrink by the mean run length and same-address atomics never pile up.
#include "skb_common.cuh"

namespace {

constexpr int INT_MAX_ = 0x7fffffff;

template <typename T> __device__ __forceinline__ void load8(const T* p, long long i, long long n, int out[8]);
template <> __device__ __forceinline__ void load8<int>(const int* p, long long i, long long n, int out[8]) {
    if (i + 8 <= n && (reinterpret_cast<uintptr_t>(p + i) & 15u) == 0) {
        const uint4 a = skb_ld_stream16(p + i), b = skb_ld_stream16(p + i + 4);
        out[0] = (int)a.x; out[1] = (int)a.y; out[2] = (int)a.z; out[3] = (int)a.w;
        out[4] = (int)b.x; out[5] = (int)b.y; out[6] = (int)b.z; out[7] = (int)b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = i + j < n ? p[i + j] : 0;
    }
}
template <> __device__ __forceinline__ void load8<short>(const short* p, long long i, long long n, int out[8]) {
    if (i + 8 <= n && (reinterpret_cast<uintptr_t>(p + i) & 15u) == 0) {
        const uint4 a = skb_ld_stream16(p + i);
        const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = (int)(short)(w[j >> 1] >> (16 * (j & 1)));
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = i + j < n ? (int)p[i + j] : 0;
    }
}

// label of the voxel before a thread's 8 (for run-start detection); -1 at the very beginning
template <typename T> __device__ __forceinline__ int prev_label(const T* p, long long i) { return i > 0 ? (int)p[i - 1] : -1; }

// ---- renumber ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) first_occurrence_kernel(const T* __restrict__ labels, long long n, int table_size,
                                                              int* __restrict__ first, unsigned* status) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    int l[8];
    load8<T>(labels, i, n, l);
    int prev = prev_label(labels, i);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (i + j < n && l[j] != prev && l[j] != 0) {
            if (l[j] < 0 || l[j] >= table_size) atomicOr(status, SKB_STATUS_LABEL_RANGE);
            else atomicMin(first + l[j], (int)(i + j));  // a later run of the same label loses against the earlier one
        }
        prev = l[j];
    }
}

// bit (first voxel) + count per 4096-voxel chunk, for every label that occurs
__global__ void __launch_bounds__(256) mark_first_kernel(const int* __restrict__ first, int table_size, ull* __restrict__ bitmap,
                                                        int* __restrict__ chunks) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= table_size) return;
    const int f = first[l];
    if (f == INT_MAX_) return;
    atomicOr(bitmap + (f >> 6), 1ull << (f & 63));
    atomicAdd(chunks + (f >> 12), 1);
}

__device__ __forceinline__ int block_exclusive_scan(int sum, int* warp_sums, int* total) {
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

// two-level exclusive scan of `vals[0..n)` in place: tiles of 8192, then (<= 4096) tile totals; *total_out = sum
constexpr int SCAN_TILE = 8192;
__global__ void __launch_bounds__(1024) scan_tiles_kernel(int* vals, long long n, int* tiles) {
    __shared__ int warp_sums[32];
    __shared__ int total;
    constexpr int PER = SCAN_TILE / 1024;
    const long long at = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * PER;
    int v[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        v[j] = at + j < n ? vals[at + j] : 0;
        sum += v[j];
    }
    int run = block_exclusive_scan(sum, warp_sums, &total);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        if (at + j < n) vals[at + j] = run;
        run += v[j];
    }
    if (threadIdx.x == 0) tiles[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) scan_top_kernel(int* tiles, long long n_tiles, int* total_out) {
    __shared__ int warp_sums[32];
    __shared__ int total;
    constexpr int PER = 4;
    const long long at = (long long)threadIdx.x * PER;
    int v[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        v[j] = at + j < n_tiles ? tiles[at + j] : 0;
        sum += v[j];
    }
    int run = block_exclusive_scan(sum, warp_sums, &total);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        if (at + j < n_tiles) tiles[at + j] = run;
        run += v[j];
    }
    if (threadIdx.x == 0 && total_out) *total_out = total;
}

// new id = 1 + number of first-occurrence voxels before mine
__global__ void __launch_bounds__(256) rank_first_kernel(const int* __restrict__ first, int table_size, const ull* __restrict__ bitmap,
                                                        const int* __restrict__ chunks, const int* __restrict__ tiles,
                                                        int* __restrict__ remap) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= table_size) return;
    const int f = first[l];
    if (f == INT_MAX_) { remap[l] = 0; return; }
    const int c = f >> 12, w = f >> 6;
    int rank = chunks[c] + tiles[c / SCAN_TILE];
    for (int j = c << 6; j < w; ++j) rank += __popcll(bitmap[j]);
    rank += __popcll(bitmap[w] & ((1ull << (f & 63)) - 1ull));
    remap[l] = rank + 1;
}

template <typename T>
__global__ void __launch_bounds__(256) apply_remap_kernel(T* __restrict__ labels, long long n, int table_size,
                                                         const int* __restrict__ remap) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    int l[8];
    load8<T>(labels, i, n, l);
    int prev = 0, prev_new = 0;
    T o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int nl = 0;
        if (l[j] > 0 && l[j] < table_size) nl = l[j] == prev ? prev_new : __ldg(remap + l[j]);
        else if (l[j] != 0) nl = l[j];  // out of range: left alone (status already says so)
        prev = l[j]; prev_new = nl;
        o[j] = (T)nl;
    }
    if (i + 8 <= n && (reinterpret_cast<uintptr_t>(labels + i) & 15u) == 0) {
        if (sizeof(T) == 4) {
            skb_st_stream16(labels + i, make_uint4((unsigned)o[0], (unsigned)o[1], (unsigned)o[2], (unsigned)o[3]));
            skb_st_stream16(labels + i + 4, make_uint4((unsigned)o[4], (unsigned)o[5], (unsigned)o[6], (unsigned)o[7]));
        } else {
            skb_st_stream16(labels + i, make_uint4(((unsigned)o[0] & 0xffffu) | ((unsigned)o[1] << 16), ((unsigned)o[2] & 0xffffu) | ((unsigned)o[3] << 16),
                                                   ((unsigned)o[4] & 0xffffu) | ((unsigned)o[5] << 16), ((unsigned)o[6] & 0xffffu) | ((unsigned)o[7] << 16)));
        }
    } else {
        for (int j = 0; j < 8 && i + j < n; ++j) labels[i + j] = o[j];
    }
}

struct RenumberLayout {
    long long n_words, n_chunks, n_tiles;
    size_t off_first, off_bitmap, off_chunks, off_tiles, total;
};
RenumberLayout renumber_layout(int64_t n, int64_t table_size) {
    RenumberLayout L;
    L.n_words = (n + 63) / 64;
    L.n_chunks = (n + 4095) / 4096;
    L.n_tiles = (L.n_chunks + SCAN_TILE - 1) / SCAN_TILE;
    size_t at = 0;
    L.off_first = at;  at = skb_align_up(at + (size_t)table_size * 4, 256);
    L.off_bitmap = at; at = skb_align_up(at + (size_t)L.n_words * 8, 256);
    L.off_chunks = at; at = skb_align_up(at + (size_t)(L.n_chunks + 1) * 4, 256);
    L.off_tiles = at;  at = skb_align_up(at + (size_t)(L.n_tiles + 1) * 4, 256);
    L.total = at;
    return L;
}

int check_labels_args(const char* who, const void* p, int dtype, int64_t n, int64_t table_size) {
    if (!p || n <= 0 || n > 2147483648LL) { skb_set_error("%s: NULL labels or voxel count outside 1..2^31", who); return SKB_E_ARG; }
    if (dtype != SKB_I16 && dtype != SKB_I32) { skb_set_error("%s: label dtype must be i16 or i32", who); return SKB_E_ARG; }
    if (table_size < 2 || table_size > (1LL << 28)) { skb_set_error("%s: table_size must be in 2..2^28 (and exceed the largest label)", who); return SKB_E_ARG; }
    return SKB_OK;
}

// ---- unique / contingency -----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) presence_kernel(const T* __restrict__ labels, long long n, int table_size,
                                                      int* __restrict__ present, unsigned* status) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    int l[8];
    load8<T>(labels, i, n, l);
    int prev = prev_label(labels, i);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (i + j < n && l[j] != prev && l[j] > 0) {  // the reference keeps `unique > 0` only (validate/lib.py:202,205)
            if (l[j] >= table_size) atomicOr(status, SKB_STATUS_LABEL_RANGE);
            else present[l[j]] = 1;
        }
        prev = l[j];
    }
}

// after the exclusive scan of `present` into `index`: index[l] = rank among the present labels, or -1
__global__ void __launch_bounds__(256) finish_index_kernel(const int* __restrict__ present, int* __restrict__ index, int table_size,
                                                          int* __restrict__ values, const int* total) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= table_size) return;
    if (present[l]) {
        if (values) values[index[l]] = l;
    } else {
        index[l] = -1;
    }
    (void)total;
}

template <typename TA, typename TB>
__global__ void __launch_bounds__(256) contingency_kernel(const TA* __restrict__ a, const TB* __restrict__ b, long long n,
                                                         const int* __restrict__ idx_a, int ta, const int* __restrict__ idx_b, int tb,
                                                         int M, int* __restrict__ inter, int* __restrict__ area_a,
                                                         int* __restrict__ area_b) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    int la[8], lb[8];
    load8<TA>(a, i, n, la);
    load8<TB>(b, i, n, lb);
    // run-length accumulate over my 8 voxels: one atomic per (run of equal a) / (run of equal b) / (run of equal pair)
    int ra = 0, ca = 0, rb = 0, cb = 0, pa = 0, pb = 0, cp = 0;
    auto flush_a = [&]() { if (ca && ra > 0 && ra < ta && idx_a[ra] >= 0) atomicAdd(area_a + idx_a[ra], ca); };
    auto flush_b = [&]() { if (cb && rb > 0 && rb < tb && idx_b[rb] >= 0) atomicAdd(area_b + idx_b[rb], cb); };
    auto flush_p = [&]() {
        if (cp && pa > 0 && pb > 0 && pa < ta && pb < tb && idx_a[pa] >= 0 && idx_b[pb] >= 0)
            atomicAdd(inter + (size_t)idx_a[pa] * M + idx_b[pb], cp);
    };
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (i + j >= n) break;
        if (la[j] != ra) { flush_a(); ra = la[j]; ca = 0; }
        if (lb[j] != rb) { flush_b(); rb = lb[j]; cb = 0; }
        if (la[j] != pa || lb[j] != pb) { flush_p(); pa = la[j]; pb = lb[j]; cp = 0; }
        ++ca; ++cb; ++cp;
    }
    flush_a(); flush_b(); flush_p();
}

__global__ void __launch_bounds__(256) iou_dice_kernel(const int* __restrict__ inter, const int* __restrict__ area_a,
                                                      const int* __restrict__ area_b, int N, int M, float* __restrict__ iou,
                                                      float* __restrict__ dice) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * M) return;
    const int i = (int)(t / M), j = (int)(t - (long long)i * M);
    const int I = inter[t];
    const long long A = area_a[i], B = area_b[j];
    // the reference divides 0-dim int64 tensors: both sides go to float32, then one fp32 division; no contact -> 0.0
    if (iou) iou[t] = I ? __fdiv_rn((float)I, (float)(A + B - I)) : 0.f;
    if (dice) dice[t] = I ? __fdiv_rn((float)(2LL * I), (float)(A + B)) : 0.f;
}

// row / column maxima of a non-negative matrix compared with thr: hit[i] = 1 if max_j iou[i,j] > thr, hit[N+j] likewise
__global__ void __launch_bounds__(256) iou_hits_kernel(const float* __restrict__ iou, int N, int M, float thr, int* __restrict__ hit) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * M) return;
    if (iou[t] > thr) {
        const int i = (int)(t / M), j = (int)(t - (long long)i * M);
        hit[i] = 1;
        hit[N + j] = 1;
    }
}
__global__ void __launch_bounds__(256) count_hits_kernel(const int* __restrict__ hit, int N, int M, int* __restrict__ out3) {
    // out3 = [true positives (gt objects hit), false positives (pred objects not hit), false negatives (gt objects not hit)]
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N + M) return;
    const int h = hit[t];
    if (t < N) atomicAdd(out3 + (h ? 0 : 2), 1);
    else if (!h) atomicAdd(out3 + 1, 1);
}

template <typename T>
__global__ void __launch_bounds__(256) label_max_kernel(const T* __restrict__ labels, long long n, int* __restrict__ out) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    int m = 0;
    if (i < n) {
        int l[8];
        load8<T>(labels, i, n, l);
#pragma unroll
        for (int j = 0; j < 8; ++j) m = max(m, l[j]);  // load8 pads with 0
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);  // one atomic per warp that saw a label
}

__global__ void __launch_bounds__(256) fill_kernel(int* p, int n, int value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = value;
}

__global__ void __launch_bounds__(256) add_tile_base_kernel(int* idx, const int* __restrict__ tiles, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] += tiles[i / SCAN_TILE];
}

}  // namespace

extern "C" size_t skb_renumber_workspace_bytes(int64_t n_voxels, int64_t table_size) {
    if (n_voxels <= 0 || table_size < 2) return 0;
    return renumber_layout(n_voxels, table_size).total;
}

extern "C" int skb_renumber(void* labels, int dtype, int64_t n_voxels, int64_t table_size, void* workspace,
                            size_t workspace_bytes, int32_t* remap, int32_t* n_labels, uint32_t* status, void* stream) {
    int rc = check_labels_args("skb_renumber", labels, dtype, n_voxels, table_size);
    if (rc) return rc;
    SKB_REQUIRE(workspace && remap && status, "skb_renumber: NULL pointer");
    const RenumberLayout L = renumber_layout(n_voxels, table_size);
    if (workspace_bytes < L.total) {
        skb_set_error("skb_renumber: workspace %zu < required %zu bytes", workspace_bytes, L.total);
        return SKB_E_WORKSPACE;
    }
    char* base = static_cast<char*>(workspace);
    int* first = reinterpret_cast<int*>(base + L.off_first);
    ull* bitmap = reinterpret_cast<ull*>(base + L.off_bitmap);
    int* chunks = reinterpret_cast<int*>(base + L.off_chunks);
    int* tiles = reinterpret_cast<int*>(base + L.off_tiles);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(base + L.off_bitmap, 0, L.off_tiles + (size_t)(L.n_tiles + 1) * 4 - L.off_bitmap, st);
    cudaMemsetAsync(status, 0, 4, st);
    const unsigned vb = (unsigned)((n_voxels + 2047) / 2048), tbk = (unsigned)((table_size + 255) / 256);
    const int T = (int)table_size;
    fill_kernel<<<tbk, 256, 0, st>>>(first, T, INT_MAX_);
    if (dtype == SKB_I32) first_occurrence_kernel<int><<<vb, 256, 0, st>>>(static_cast<const int*>(labels), n_voxels, T, first, status);
    else first_occurrence_kernel<short><<<vb, 256, 0, st>>>(static_cast<const short*>(labels), n_voxels, T, first, status);
    mark_first_kernel<<<tbk, 256, 0, st>>>(first, T, bitmap, chunks);
    scan_tiles_kernel<<<(unsigned)L.n_tiles, 1024, 0, st>>>(chunks, L.n_chunks, tiles);
    scan_top_kernel<<<1, 1024, 0, st>>>(tiles, L.n_tiles, n_labels);
    rank_first_kernel<<<tbk, 256, 0, st>>>(first, T, bitmap, chunks, tiles, remap);
    if (dtype == SKB_I32) apply_remap_kernel<int><<<vb, 256, 0, st>>>(static_cast<int*>(labels), n_voxels, T, remap);
    else apply_remap_kernel<short><<<vb, 256, 0, st>>>(static_cast<short*>(labels), n_voxels, T, remap);
    SKB_LAUNCH_CHECK("skb_renumber");
    return SKB_OK;
}

extern "C" int skb_unique_index(const void* labels, int dtype, int64_t n_voxels, int64_t table_size, int32_t* index,
                                int32_t* values, int32_t* count, void* workspace, size_t workspace_bytes, uint32_t* status,
                                void* stream) {
    int rc = check_labels_args("skb_unique_index", labels, dtype, n_voxels, table_size);
    if (rc) return rc;
    SKB_REQUIRE(index && count && workspace && status, "skb_unique_index: NULL pointer");
    const long long n_tiles = (table_size + SCAN_TILE - 1) / SCAN_TILE;
    SKB_REQUIRE(n_tiles <= 4096, "skb_unique_index: table too large");
    const size_t need = skb_align_up((size_t)table_size * 4, 256) + (size_t)(n_tiles + 1) * 4;
    if (workspace_bytes < need) {
        skb_set_error("skb_unique_index: workspace %zu < required %zu bytes", workspace_bytes, need);
        return SKB_E_WORKSPACE;
    }
    int* present = static_cast<int*>(workspace);
    int* tiles = reinterpret_cast<int*>(static_cast<char*>(workspace) + skb_align_up((size_t)table_size * 4, 256));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(present, 0, (size_t)table_size * 4, st);
    const unsigned vb = (unsigned)((n_voxels + 2047) / 2048), tbk = (unsigned)((table_size + 255) / 256);
    const int T = (int)table_size;
    if (dtype == SKB_I32) presence_kernel<int><<<vb, 256, 0, st>>>(static_cast<const int*>(labels), n_voxels, T, present, status);
    else presence_kernel<short><<<vb, 256, 0, st>>>(static_cast<const short*>(labels), n_voxels, T, present, status);
    cudaMemcpyAsync(index, present, (size_t)table_size * 4, cudaMemcpyDeviceToDevice, st);
    scan_tiles_kernel<<<(unsigned)n_tiles, 1024, 0, st>>>(index, table_size, tiles);
    scan_top_kernel<<<1, 1024, 0, st>>>(tiles, n_tiles, count);
    add_tile_base_kernel<<<tbk, 256, 0, st>>>(index, tiles, T);  // second scan level folded in
    finish_index_kernel<<<tbk, 256, 0, st>>>(present, index, T, values, count);
    SKB_LAUNCH_CHECK("skb_unique_index");
    return SKB_OK;
}

extern "C" size_t skb_unique_index_workspace_bytes(int64_t table_size) {
    if (table_size < 2) return 0;
    const long long n_tiles = (table_size + SCAN_TILE - 1) / SCAN_TILE;
    return skb_align_up((size_t)table_size * 4, 256) + (size_t)(n_tiles + 1) * 4;
}

extern "C" int skb_contingency(const void* gt, int gt_dtype, const void* pred, int pred_dtype, int64_t n_voxels,
                               const int32_t* index_gt, int64_t table_gt, const int32_t* index_pred, int64_t table_pred,
                               int64_t N, int64_t M, int32_t* inter_zeroed, int32_t* area_gt_zeroed,
                               int32_t* area_pred_zeroed, void* stream) {
    int rc = check_labels_args("skb_contingency", gt, gt_dtype, n_voxels, table_gt);
    if (rc) return rc;
    rc = check_labels_args("skb_contingency", pred, pred_dtype, n_voxels, table_pred);
    if (rc) return rc;
    SKB_REQUIRE(index_gt && index_pred && N >= 0 && M >= 0 && N * M < (1LL << 31), "skb_contingency: bad tables (N*M must stay below 2^31)");
    if (N == 0 || M == 0) return SKB_OK;
    SKB_REQUIRE(inter_zeroed && area_gt_zeroed && area_pred_zeroed, "skb_contingency: NULL output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned vb = (unsigned)((n_voxels + 2047) / 2048);
    const int ta = (int)table_gt, tb = (int)table_pred, m = (int)M;
#define SKB_CONT(TA, TB)                                                                                                 \
    contingency_kernel<TA, TB><<<vb, 256, 0, st>>>(static_cast<const TA*>(gt), static_cast<const TB*>(pred), n_voxels,   \
                                                   index_gt, ta, index_pred, tb, m, inter_zeroed, area_gt_zeroed, area_pred_zeroed)
    if (gt_dtype == SKB_I32 && pred_dtype == SKB_I32) SKB_CONT(int, int);
    else if (gt_dtype == SKB_I32) SKB_CONT(int, short);
    else if (pred_dtype == SKB_I32) SKB_CONT(short, int);
    else SKB_CONT(short, short);
#undef SKB_CONT
    SKB_LAUNCH_CHECK("contingency_kernel");
    return SKB_OK;
}

extern "C" int skb_iou_dice(const int32_t* inter, const int32_t* area_gt, const int32_t* area_pred, int64_t N, int64_t M,
                            float* iou, float* dice, void* stream) {
    SKB_REQUIRE(N >= 0 && M >= 0 && N * M < (1LL << 31), "skb_iou_dice: bad shape");
    if (N == 0 || M == 0) return SKB_OK;
    SKB_REQUIRE(inter && area_gt && area_pred && (iou || dice), "skb_iou_dice: NULL pointer");
    const long long total = N * M;
    iou_dice_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(inter, area_gt, area_pred, (int)N,
                                                                                                 (int)M, iou, dice);
    SKB_LAUNCH_CHECK("iou_dice_kernel");
    return SKB_OK;
}

extern "C" int skb_accuracies_from_iou(const float* iou, int64_t N, int64_t M, float thr, int32_t* hits_scratch,
                                       int32_t* out3, void* stream) {
    SKB_REQUIRE(N > 0 && M > 0 && N * M < (1LL << 31), "skb_accuracies_from_iou: needs a non-empty matrix");
    SKB_REQUIRE(iou && hits_scratch && out3, "skb_accuracies_from_iou: NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(hits_scratch, 0, (size_t)(N + M) * 4, st);
    cudaMemsetAsync(out3, 0, 12, st);
    const long long total = N * M;
    iou_hits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(iou, (int)N, (int)M, thr, hits_scratch);
    count_hits_kernel<<<(unsigned)((N + M + 255) / 256), 256, 0, st>>>(hits_scratch, (int)N, (int)M, out3);
    SKB_LAUNCH_CHECK("skb_accuracies_from_iou");
    return SKB_OK;
}

extern "C" int skb_label_max(const void* labels, int dtype, int64_t n_voxels, int32_t* max_out, void* stream) {
    int rc = check_labels_args("skb_label_max", labels, dtype, n_voxels, 2);
    if (rc) return rc;
    SKB_REQUIRE(max_out, "skb_label_max: NULL output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(max_out, 0, 4, st);
    const unsigned vb = (unsigned)((n_voxels + 2047) / 2048);
    if (dtype == SKB_I32) label_max_kernel<int><<<vb, 256, 0, st>>>(static_cast<const int*>(labels), n_voxels, max_out);
    else label_max_kernel<short><<<vb, 256, 0, st>>>(static_cast<const short*>(labels), n_voxels, max_out);
    SKB_LAUNCH_CHECK("label_max_kernel");
    return SKB_OK;
}

/* labels[i] = table[labels[i]] for 0 < labels[i] < table_size, in place (the `replace` step of the reference's
 * efficient_flood_fill, flood_fill.py:206-234, as one streaming pass with a look-up table) */
extern "C" int skb_apply_label_table(void* labels, int dtype, int64_t n_voxels, const int32_t* table, int64_t table_size,
                                     void* stream) {
    int rc = check_labels_args("skb_apply_label_table", labels, dtype, n_voxels, table_size);
    if (rc) return rc;
    SKB_REQUIRE(table, "skb_apply_label_table: NULL table");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned vb = (unsigned)((n_voxels + 2047) / 2048);
    if (dtype == SKB_I32) apply_remap_kernel<int><<<vb, 256, 0, st>>>(static_cast<int*>(labels), n_voxels, (int)table_size, table);
    else apply_remap_kernel<short><<<vb, 256, 0, st>>>(static_cast<short*>(labels), n_voxels, (int)table_size, table);
    SKB_LAUNCH_CHECK("apply_remap_kernel");
    return SKB_OK;
}

