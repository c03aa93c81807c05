// Training-side kernels of the path (kernel (c) of the north star and friends).
//
//   skb_embed_prob_fwd/_bwd   skoots/lib/embedding_to_prob.py:5-51
//                             p = exp( sum_c (E_c - S_c)^2 / (-2 (sigma_c + eps)^2) )
//   skb_vec_prob_fwd/_bwd     the same with E = idx + v*s computed on the fly from the network's
//                             vector head (vector_to_embedding N=1 fused in; train/engine.py:465-466),
//                             so the fp32 embedding never goes to HBM.
//   skb_bake_skeleton         skoots/lib/skeleton.py:370-445 (CPU/torch semantics): for every voxel of
//                             object k the nearest point of skeleton k, anisotropy scaling the
//                             coordinates, first minimum wins.  The point table is staged into shared
//                             memory with one TMA bulk copy (cp.async.bulk + mbarrier); the search
//                             is an fp32 min-reduction per voxel — no tensor cores, by design.
//   skb_stamp_disks           skoots/lib/skeleton.py:531-593 (skeleton_to_mask)
#include "skb_common.cuh"

// ------------------------------------------------------------------------------------------
// embedding -> probability
// ------------------------------------------------------------------------------------------
struct ProbParams {
    int C;               // 2 or 3
    long long inner;     // voxels per (b,c) plane
    long long total;     // B * inner
    float neg2sig2[3];   // -2 (sigma+eps)^2, rounded like the reference (fp32 ops)
    float scale[3];      // fused form only
    int X, Y, Z;         // fused form only (Z = 1, Y = last dim for 2-D)
};

template <typename ST>
__global__ void __launch_bounds__(256) embed_prob_fwd_kernel(const float* __restrict__ E, const ST* __restrict__ S,
                                                            float* __restrict__ out, ProbParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.total) return;
    const long long b = i / P.inner, r = i - b * P.inner;
    const long long base = b * P.C * P.inner + r;
    float acc = 0.f;
#pragma unroll 3
    for (int c = 0; c < P.C; ++c) {
        const float d = __fsub_rn(E[base + c * P.inner], skb_to_float<ST>(S[base + c * P.inner]));
        acc = __fadd_rn(acc, __fdiv_rn(__fmul_rn(d, d), P.neg2sig2[c]));
    }
    out[i] = expf(acc);
}

// grad wrt embedding (fp32) and optionally wrt the baked skeleton (same dtype as S)
template <typename ST>
__global__ void __launch_bounds__(256) embed_prob_bwd_kernel(const float* __restrict__ E, const ST* __restrict__ S,
                                                            const float* __restrict__ prob, const float* __restrict__ go,
                                                            float* __restrict__ gE, ST* __restrict__ gS, ProbParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.total) return;
    const long long b = i / P.inner, r = i - b * P.inner;
    const long long base = b * P.C * P.inner + r;
    const float gp = go[i] * prob[i];
    for (int c = 0; c < P.C; ++c) {
        const float d = E[base + c * P.inner] - skb_to_float<ST>(S[base + c * P.inner]);
        const float g = gp * 2.f * d / P.neg2sig2[c];
        if (gE) gE[base + c * P.inner] = g;
        if (gS) gS[base + c * P.inner] = skb_from_float<ST>(-g);
    }
}

template <typename VT, typename ST, bool BWD>
__global__ void __launch_bounds__(256) vec_prob_kernel(const VT* __restrict__ vec, const ST* __restrict__ S,
                                                      float* __restrict__ prob, const float* __restrict__ go,
                                                      VT* __restrict__ gvec, ProbParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.total) return;
    const long long b = i / P.inner, r = i - b * P.inner;
    const long long base = b * P.C * P.inner + r;
    int idx[3];
    if (P.C == 3) {
        idx[2] = (int)(r % P.Z);
        const long long q = r / P.Z;
        idx[1] = (int)(q % P.Y);
        idx[0] = (int)(q / P.Y);
    } else {
        idx[1] = (int)(r % P.Y);
        idx[0] = (int)(r / P.Y);
        idx[2] = 0;
    }
    float d[3], acc = 0.f;
    for (int c = 0; c < P.C; ++c) {
        const float e = __fadd_rn((float)idx[c], __fmul_rn(skb_to_float<VT>(vec[base + c * P.inner]), P.scale[c]));
        d[c] = __fsub_rn(e, skb_to_float<ST>(S[base + c * P.inner]));
        acc = __fadd_rn(acc, __fdiv_rn(__fmul_rn(d[c], d[c]), P.neg2sig2[c]));
    }
    const float p = expf(acc);
    if (!BWD) {
        prob[i] = p;
    } else {
        const float gp = go[i] * p;
        for (int c = 0; c < P.C; ++c)
            gvec[base + c * P.inner] = skb_from_float<VT>(gp * 2.f * d[c] / P.neg2sig2[c] * P.scale[c]);
    }
}

static int fill_prob(ProbParams& P, int64_t B, int C, int64_t inner, const float* sigma, float eps) {
    SKB_REQUIRE(B >= 1 && (C == 2 || C == 3) && inner >= 1 && sigma, "embed_prob: bad argument");
    P.C = C; P.inner = inner; P.total = B * inner;
    for (int c = 0; c < C; ++c) {
        // sigma + eps ; pow(2) ; mul(2) ; mul(-1) — each an fp32 op in the reference (:37-38)
        volatile float s = sigma[c] + eps;
        volatile float s2 = s * s;
        volatile float s3 = s2 * 2.f;
        P.neg2sig2[c] = -s3;
    }
    return SKB_OK;
}

#define DISPATCH_S(dtype, FN)                                              \
    if (dtype == SKB_F32) { using ST = float; FN; }                        \
    else if (dtype == SKB_F16) { using ST = __half; FN; }                  \
    else if (dtype == SKB_BF16) { using ST = __nv_bfloat16; FN; }          \
    else { skb_set_error("unsupported dtype %d", dtype); return SKB_E_ARG; }

extern "C" int skb_embed_prob_fwd(const float* embedding, const void* baked, int baked_dtype, int64_t B, int C,
                                  int64_t inner, const float* sigma, float eps, float* out, void* stream) {
    ProbParams P = {};
    int rc = fill_prob(P, B, C, inner, sigma, eps);
    if (rc) return rc;
    SKB_REQUIRE(embedding && baked && out, "skb_embed_prob_fwd: NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned nb = (unsigned)((P.total + 255) / 256);
    DISPATCH_S(baked_dtype, (embed_prob_fwd_kernel<ST><<<nb, 256, 0, st>>>(embedding, static_cast<const ST*>(baked), out, P)));
    SKB_LAUNCH_CHECK("embed_prob_fwd_kernel");
    return SKB_OK;
}

extern "C" int skb_embed_prob_bwd(const float* embedding, const void* baked, int baked_dtype, const float* prob,
                                  const float* grad_out, int64_t B, int C, int64_t inner, const float* sigma, float eps,
                                  float* grad_embedding, void* grad_baked, void* stream) {
    ProbParams P = {};
    int rc = fill_prob(P, B, C, inner, sigma, eps);
    if (rc) return rc;
    SKB_REQUIRE(embedding && baked && prob && grad_out && (grad_embedding || grad_baked), "skb_embed_prob_bwd: NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned nb = (unsigned)((P.total + 255) / 256);
    DISPATCH_S(baked_dtype, (embed_prob_bwd_kernel<ST><<<nb, 256, 0, st>>>(embedding, static_cast<const ST*>(baked), prob,
                                                                          grad_out, grad_embedding,
                                                                          static_cast<ST*>(grad_baked), P)));
    SKB_LAUNCH_CHECK("embed_prob_bwd_kernel");
    return SKB_OK;
}

template <typename VT>
static int launch_vec_prob(const void* vec, const void* baked, int baked_dtype, float* prob, const float* go, void* gvec,
                           const ProbParams& P, cudaStream_t st) {
    const unsigned nb = (unsigned)((P.total + 255) / 256);
    const VT* v = static_cast<const VT*>(vec);
    VT* gv = static_cast<VT*>(gvec);
    if (go) {
        DISPATCH_S(baked_dtype, (vec_prob_kernel<VT, ST, true><<<nb, 256, 0, st>>>(v, static_cast<const ST*>(baked), prob, go, gv, P)));
    } else {
        DISPATCH_S(baked_dtype, (vec_prob_kernel<VT, ST, false><<<nb, 256, 0, st>>>(v, static_cast<const ST*>(baked), prob, go, gv, P)));
    }
    return SKB_OK;
}

// fused vector_to_embedding(N=1) + baked_embed_to_prob.  grad_out == NULL: forward (writes prob);
// otherwise backward (writes grad_vec, recomputing the probability instead of re-reading it).
extern "C" int skb_vec_prob(const void* vec, int vec_dtype, const void* baked, int baked_dtype, int64_t B, int C,
                            int64_t X, int64_t Y, int64_t Z, const float* scale, const float* sigma, float eps,
                            float* prob, const float* grad_out, void* grad_vec, void* stream) {
    ProbParams P = {};
    const int64_t inner = C == 3 ? X * Y * Z : X * Y;
    int rc = fill_prob(P, B, C, inner, sigma, eps);
    if (rc) return rc;
    SKB_REQUIRE(vec && baked && scale && (grad_out ? grad_vec != nullptr : prob != nullptr), "skb_vec_prob: NULL pointer");
    SKB_REQUIRE(X < (1 << 24) && Y < (1 << 24) && Z < (1 << 24), "skb_vec_prob: axis too long");
    for (int c = 0; c < C; ++c) P.scale[c] = scale[c];
    P.X = (int)X; P.Y = (int)Y; P.Z = (int)(C == 3 ? Z : 1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec_dtype == SKB_F32) rc = launch_vec_prob<float>(vec, baked, baked_dtype, prob, grad_out, grad_vec, P, st);
    else if (vec_dtype == SKB_F16) rc = launch_vec_prob<__half>(vec, baked, baked_dtype, prob, grad_out, grad_vec, P, st);
    else if (vec_dtype == SKB_BF16) rc = launch_vec_prob<__nv_bfloat16>(vec, baked, baked_dtype, prob, grad_out, grad_vec, P, st);
    else SKB_REQUIRE(false, "skb_vec_prob: vec dtype");
    if (rc) return rc;
    SKB_LAUNCH_CHECK("vec_prob_kernel");
    return SKB_OK;
}

// ------------------------------------------------------------------------------------------
// bake_skeleton: nearest skeleton point per voxel (min-reduction; TMA-staged point table)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// one elected thread: stage `bytes` (multiple of 16) from global to shared with a TMA bulk copy
__device__ __forceinline__ void tma_stage(void* smem_dst, const void* gsrc, unsigned bytes, ull* bar) {
    const unsigned b = smem_u32(bar), d = smem_u32(smem_dst);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                 "l"(gsrc), "r"(bytes), "r"(b)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(ull* bar, unsigned phase) {
    const unsigned b = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b),
        "r"(phase)
        : "memory");
}

struct BakeParams {
    int X, Y, Z;
    int n_ids;
    int n_points;
    int staged_points;   // how many points fit the shared-memory stage
    float an[3];
    const int* ids;      // sorted object ids
    const int* offsets;  // n_ids + 1 prefix of point counts
    const float* points; // (n_points, 4) xyz + pad, 16-byte rows
    unsigned* status;    // bit 1: a mask id has no skeleton (reference raises KeyError)
};

template <typename MT>
__global__ void __launch_bounds__(256) bake_kernel(const MT* __restrict__ mask, float* __restrict__ baked,
                                                  float* __restrict__ dist_out, BakeParams P, long long V) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* s_pts = reinterpret_cast<float4*>(smem_raw);
    __shared__ __align__(8) ull bar;
    if (threadIdx.x == 0 && P.staged_points > 0) tma_stage(s_pts, P.points, (unsigned)P.staged_points * 16u, &bar);
    __syncthreads();  // barrier init visible to everyone before they wait on it

    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int id = 0;
    int x = 0, y = 0, z = 0;
    if (i < V) {
        id = (int)mask[i];
        z = (int)(i % P.Z);
        const long long q = i / P.Z;
        y = (int)(q % P.Y);
        x = (int)(q / P.Y);
    }
    int lo = 0, hi = 0;
    if (id != 0) {
        int a = 0, b = P.n_ids - 1, found = -1;
        while (a <= b) {
            int m = (a + b) >> 1, v = __ldg(P.ids + m);
            if (v == id) { found = m; break; }
            if (v < id) a = m + 1; else b = m - 1;
        }
        if (found < 0) atomicOr(P.status, 2u);
        else { lo = __ldg(P.offsets + found); hi = __ldg(P.offsets + found + 1); }
    }
    if (P.staged_points > 0) mbar_wait(&bar, 0);

    if (i >= V) return;
    float bx = 0.f, by = 0.f, bz = 0.f, best = INFINITY;
    const float ax = __fmul_rn(P.an[0], (float)x), ay = __fmul_rn(P.an[1], (float)y), az = __fmul_rn(P.an[2], (float)z);
    for (int k = lo; k < hi; ++k) {
        const float4 p = k < P.staged_points ? s_pts[k] : __ldg(reinterpret_cast<const float4*>(P.points) + k);
        const float dx = __fsub_rn(__fmul_rn(p.x, P.an[0]), ax);
        const float dy = __fsub_rn(__fmul_rn(p.y, P.an[1]), ay);
        const float dz = __fsub_rn(__fmul_rn(p.z, P.an[2]), az);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const float d = sqrtf(d2);  // the reference compares cdist's sqrt'ed values (first argmin)
        if (d < best) { best = d; bx = p.x; by = p.y; bz = p.z; }
    }
    baked[i] = bx;
    baked[i + V] = by;
    baked[i + 2 * V] = bz;
    if (dist_out) dist_out[i] = (lo < hi) ? best : 0.f;
}

extern "C" int skb_bake_skeleton(const void* mask, int mask_dtype, int64_t X, int64_t Y, int64_t Z, const int32_t* ids,
                                 const int32_t* offsets, int n_ids, const float* points_xyzw, int n_points,
                                 const float anisotropy[3], float* baked, float* distance, uint32_t* status,
                                 void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_bake_skeleton");
    if (rc) return rc;
    SKB_REQUIRE(mask && baked && status && anisotropy && n_ids >= 0 && n_points >= 0, "skb_bake_skeleton: bad argument");
    SKB_REQUIRE(n_ids == 0 || (ids && offsets && points_xyzw), "skb_bake_skeleton: NULL tables");
    SKB_REQUIRE(n_points == 0 || skb_aligned16(points_xyzw), "skb_bake_skeleton: point table must be 16-byte aligned");
    BakeParams P;
    P.X = (int)X; P.Y = (int)Y; P.Z = (int)Z;
    P.n_ids = n_ids; P.n_points = n_points;
    const int max_stage = (160 * 1024) / 16;  // 160 KB of the 227 KB: leaves room for two CTAs of small tables
    P.staged_points = n_points < max_stage ? n_points : max_stage;
    P.an[0] = anisotropy[0]; P.an[1] = anisotropy[1]; P.an[2] = anisotropy[2];
    P.ids = ids; P.offsets = offsets; P.points = points_xyzw; P.status = status;
    const long long V = X * Y * Z;
    const size_t smem = (size_t)P.staged_points * 16;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(status, 0, sizeof(uint32_t), st);
    const unsigned nb = (unsigned)((V + 255) / 256);
#define BAKE_LAUNCH(MT)                                                                                             \
    do {                                                                                                            \
        cudaFuncSetAttribute(bake_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
        bake_kernel<MT><<<nb, 256, smem, st>>>(static_cast<const MT*>(mask), baked, distance, P, V);                \
    } while (0)
    if (mask_dtype == SKB_I32) BAKE_LAUNCH(int32_t);
    else if (mask_dtype == SKB_I16) BAKE_LAUNCH(int16_t);
    else if (mask_dtype == SKB_U8) BAKE_LAUNCH(uint8_t);
    else SKB_REQUIRE(false, "skb_bake_skeleton: mask dtype must be u8, i16 or i32");
    SKB_LAUNCH_CHECK("bake_kernel");
    return SKB_OK;
}

// ------------------------------------------------------------------------------------------
// skeleton_to_mask: OR-stamp the disk offsets around every skeleton point
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stamp_kernel(const float* __restrict__ points, int n_points,
                                                   const int* __restrict__ offsets, int n_offsets, int X, int Y, int Z,
                                                   float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_points * n_offsets) return;
    const int pi = (int)(i / n_offsets), oi = (int)(i - (long long)pi * n_offsets);
    // float point + int64 offset promotes to fp32, then .long() truncates toward zero (skeleton.py:563-569)
    const long long px = (long long)__fadd_rn(points[3 * pi + 0], (float)offsets[3 * oi + 0]);
    const long long py = (long long)__fadd_rn(points[3 * pi + 1], (float)offsets[3 * oi + 1]);
    const long long pz = (long long)__fadd_rn(points[3 * pi + 2], (float)offsets[3 * oi + 2]);
    if (px < 0 || px >= X || py < 0 || py >= Y || pz < 0 || pz >= Z) return;
    out[(px * Y + py) * Z + pz] = 1.0f;
}

extern "C" int skb_stamp_disks(const float* points_xyz, int n_points, const int32_t* offsets_xyz, int n_offsets,
                               int64_t X, int64_t Y, int64_t Z, float* out_zeroed, void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_stamp_disks");
    if (rc) return rc;
    SKB_REQUIRE(out_zeroed && n_points >= 0 && n_offsets >= 0, "skb_stamp_disks: bad argument");
    if (n_points == 0 || n_offsets == 0) return SKB_OK;
    SKB_REQUIRE(points_xyz && offsets_xyz, "skb_stamp_disks: NULL tables");
    const long long total = (long long)n_points * n_offsets;
    stamp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        points_xyz, n_points, offsets_xyz, n_offsets, (int)X, (int)Y, (int)Z, out_zeroed);
    SKB_LAUNCH_CHECK("stamp_kernel");
    return SKB_OK;
}
