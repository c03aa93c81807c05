// Training-side kernels of the path (kernel (c) of the north star and friends).
//
//   skb_embed_prob_fwd/_bwd   skoots/lib/embedding_to_prob.py:5-51
//                             p = exp( sum_c (E_c - S_c)^2 / (-2 (sigma_c + eps)^2) )
//   skb_vec_prob_fwd/_bwd     the same with E = idx + v*s computed on the fly from the network's
//                             vector head (vector_to_embedding N=1 fused in; train/engine.py:465-466),
//                             so the fp32 embedding never goes to HBM.
//   skb_bake_skeletons        skoots/lib/skeleton.py:370-445 (CPU/torch semantics) + :18-48: for every voxel of
//                             object k the nearest point of skeleton k, anisotropy scaling the
//                             coordinates, first minimum wins; then the masked 3x3x3 mean.  A CTA stages only
//                             the skeletons of the ids inside its tile into shared memory (one TMA bulk copy
//                             per id, cp.async.bulk + mbarrier); the search is an fp32 min-reduction per
//                             voxel — no tensor cores, by design.  One launch per batch.
//   skb_stamp_disks           skoots/lib/skeleton.py:531-593 (skeleton_to_mask)
#include <climits>
#include <initializer_list>

#include "skb_common.cuh"

// ------------------------------------------------------------------------------------------
// embedding -> probability
// ------------------------------------------------------------------------------------------
struct ProbParams {
    int C;               // 2 or 3
    long long inner;     // voxels per (b,c) plane
    long long total;     // B * inner
    float neg2sig2[3];   // -2 (sigma+eps)^2, rounded like the reference (fp32 ops)
    float inv[3];        // 1 / neg2sig2: the BACKWARD kernels multiply (gradients carry a tolerance, and an IEEE division is
                         // ~12 instructions: six of them per voxel made the fused backward instruction-bound, 172 us for C4)
    float scale[3];      // fused form only
    int X, Y, Z;         // fused form only (Z = 1, Y = last dim for 2-D)
};

// ---- 8 consecutive elements per thread: 16-byte streaming loads / stores --------------------------------
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&o)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&o)[8]) {
    const uint4 a = skb_ld_stream16(p), b = skb_ld_stream16(p + 4);
    o[0] = __uint_as_float(a.x); o[1] = __uint_as_float(a.y); o[2] = __uint_as_float(a.z); o[3] = __uint_as_float(a.w);
    o[4] = __uint_as_float(b.x); o[5] = __uint_as_float(b.y); o[6] = __uint_as_float(b.z); o[7] = __uint_as_float(b.w);
}
template <> __device__ __forceinline__ void ld8<__half>(const __half* p, float (&o)[8]) {
    const uint4 a = skb_ld_stream16(p);
    const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        o[2 * k] = __half2float(__ushort_as_half((unsigned short)(w[k] & 0xffffu)));
        o[2 * k + 1] = __half2float(__ushort_as_half((unsigned short)(w[k] >> 16)));
    }
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[8]) {
    const uint4 a = skb_ld_stream16(p);
    const unsigned w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        o[2 * k] = __uint_as_float(w[k] << 16);
        o[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
    skb_st_stream16(p, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])));
    skb_st_stream16(p + 4, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
}
template <> __device__ __forceinline__ void st8<__half>(__half* p, const float (&v)[8]) {
    unsigned w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        w[k] = (unsigned)__half_as_ushort(__float2half_rn(v[2 * k])) | ((unsigned)__half_as_ushort(__float2half_rn(v[2 * k + 1])) << 16);
    skb_st_stream16(p, make_uint4(w[0], w[1], w[2], w[3]));
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
    unsigned w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        w[k] = (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * k])) |
               ((unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * k + 1])) << 16);
    skb_st_stream16(p, make_uint4(w[0], w[1], w[2], w[3]));
}

// One thread = 8 consecutive voxels of one batch entry (inner % 8 == 0, every plane 16-byte aligned): the C planes of E
// and S come in as 16-byte loads, the probability leaves as two 16-byte stores.  Same fp32 operation sequence as the
// scalar kernels below, which remain for ragged shapes.
template <typename ST, int NC>
__global__ void __launch_bounds__(256) embed_prob_fwd_vec8_kernel(const float* __restrict__ E, const ST* __restrict__ S,
                                                                 float* __restrict__ out, ProbParams P, long long gpb,
                                                                 long long groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const long long b = g / gpb, r = (g - b * gpb) * 8;
    const long long base = b * NC * P.inner + r;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float e[8], s[8];
        ld8<float>(E + base + c * P.inner, e);
        ld8<ST>(S + base + c * P.inner, s);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = __fsub_rn(e[j], s[j]);
            acc[j] = __fadd_rn(acc[j], __fdiv_rn(__fmul_rn(d, d), P.neg2sig2[c]));
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = expf(acc[j]);
    st8<float>(out + b * P.inner + r, acc);
}

template <typename ST, int NC>
__global__ void __launch_bounds__(256) embed_prob_bwd_vec8_kernel(const float* __restrict__ E, const ST* __restrict__ S,
                                                                 const float* __restrict__ prob, const float* __restrict__ go,
                                                                 float* __restrict__ gE, ST* __restrict__ gS, ProbParams P,
                                                                 long long gpb, long long groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const long long b = g / gpb, r = (g - b * gpb) * 8;
    const long long base = b * NC * P.inner + r;
    float gp[8], p8[8];
    ld8<float>(go + b * P.inner + r, gp);
    ld8<float>(prob + b * P.inner + r, p8);
#pragma unroll
    for (int j = 0; j < 8; ++j) gp[j] = gp[j] * p8[j];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float e[8], s[8], ge[8], gs[8];
        ld8<float>(E + base + c * P.inner, e);
        ld8<ST>(S + base + c * P.inner, s);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = e[j] - s[j];
            ge[j] = gp[j] * 2.f * d * P.inv[c];
            gs[j] = -ge[j];
        }
        if (gE) st8<float>(gE + base + c * P.inner, ge);
        if (gS) st8<ST>(gS + base + c * P.inner, gs);
    }
}

template <typename VT, typename ST, bool BWD, int NC>
__global__ void __launch_bounds__(256) vec_prob_vec8_kernel(const VT* __restrict__ vec, const ST* __restrict__ S,
                                                           float* __restrict__ prob, const float* __restrict__ go,
                                                           VT* __restrict__ gvec, ProbParams P, long long gpb, long long groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const long long b = g / gpb, r = (g - b * gpb) * 8;
    const long long base = b * NC * P.inner + r;
    // coordinates of the 8 voxels: decompose the first, carry for the rest (a group may run over a row end: Z = 20)
    int c0[8], c1[8], c2[8];
    {
        const unsigned ur = (unsigned)r;  // inner < 2^31 (checked by the host)
        int x, y, z;
        if (NC == 3) {
            const unsigned q = ur / (unsigned)P.Z;
            z = (int)(ur - q * (unsigned)P.Z);
            x = (int)(q / (unsigned)P.Y);
            y = (int)(q - (unsigned)x * (unsigned)P.Y);
        } else {
            x = (int)(ur / (unsigned)P.Y);
            y = (int)(ur - (unsigned)x * (unsigned)P.Y);
            z = 0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            c0[j] = x; c1[j] = y; c2[j] = z;
            if (NC == 3) {
                if (++z == P.Z) { z = 0; if (++y == P.Y) { y = 0; ++x; } }
            } else {
                if (++y == P.Y) { y = 0; ++x; }
            }
        }
    }
    float d[3][8], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float v[8], s[8];
        ld8<VT>(vec + base + c * P.inner, v);
        ld8<ST>(S + base + c * P.inner, s);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int idx = c == 0 ? c0[j] : (c == 1 ? c1[j] : c2[j]);
            const float e = __fadd_rn((float)idx, __fmul_rn(v[j], P.scale[c]));
            d[c][j] = __fsub_rn(e, s[j]);
            acc[j] = BWD ? acc[j] + d[c][j] * d[c][j] * P.inv[c]
                         : __fadd_rn(acc[j], __fdiv_rn(__fmul_rn(d[c][j], d[c][j]), P.neg2sig2[c]));
        }
    }
    float p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = expf(acc[j]);
    if (!BWD) {
        st8<float>(prob + b * P.inner + r, p);
    } else {
        float gp[8];
        ld8<float>(go + b * P.inner + r, gp);
#pragma unroll
        for (int j = 0; j < 8; ++j) gp[j] = gp[j] * p[j];
    #pragma unroll
    for (int c = 0; c < NC; ++c) {
            float gv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) gv[j] = gp[j] * 2.f * d[c][j] * P.inv[c] * P.scale[c];
            st8<VT>(gvec + base + c * P.inner, gv);
        }
    }
}

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
        P.inv[c] = 1.0f / P.neg2sig2[c];
    }
    return SKB_OK;
}

#define DISPATCH_S(dtype, FN)                                              \
    if (dtype == SKB_F32) { using ST = float; FN; }                        \
    else if (dtype == SKB_F16) { using ST = __half; FN; }                  \
    else if (dtype == SKB_BF16) { using ST = __nv_bfloat16; FN; }          \
    else { skb_set_error("unsupported dtype %d", dtype); return SKB_E_ARG; }

// the 8-per-thread kernels need whole groups per plane and 16-byte aligned planes
static bool vec8_ok(int64_t inner, std::initializer_list<const void*> ptrs) {
    if (inner % 8 != 0 || inner >= (1LL << 31)) return false;
    for (const void* p : ptrs)
        if (p && !skb_aligned16(p)) return false;
    return true;
}

extern "C" int skb_embed_prob_fwd(const float* embedding, const void* baked, int baked_dtype, int64_t B, int C,
                                  int64_t inner, const float* sigma, float eps, float* out, void* stream) {
    ProbParams P = {};
    int rc = fill_prob(P, B, C, inner, sigma, eps);
    if (rc) return rc;
    SKB_REQUIRE(embedding && baked && out, "skb_embed_prob_fwd: NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec8_ok(inner, {embedding, baked, out})) {
        const long long gpb = inner / 8, groups = B * gpb;
        const unsigned nb = (unsigned)((groups + 255) / 256);
        if (C == 3) { DISPATCH_S(baked_dtype, (embed_prob_fwd_vec8_kernel<ST, 3><<<nb, 256, 0, st>>>(embedding, static_cast<const ST*>(baked), out, P, gpb, groups))); }
        else { DISPATCH_S(baked_dtype, (embed_prob_fwd_vec8_kernel<ST, 2><<<nb, 256, 0, st>>>(embedding, static_cast<const ST*>(baked), out, P, gpb, groups))); }
    } else {
        const unsigned nb = (unsigned)((P.total + 255) / 256);
        DISPATCH_S(baked_dtype, (embed_prob_fwd_kernel<ST><<<nb, 256, 0, st>>>(embedding, static_cast<const ST*>(baked), out, P)));
    }
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
    if (vec8_ok(inner, {embedding, baked, prob, grad_out, grad_embedding, grad_baked})) {
        const long long gpb = inner / 8, groups = B * gpb;
        const unsigned nb = (unsigned)((groups + 255) / 256);
        if (C == 3) { DISPATCH_S(baked_dtype, (embed_prob_bwd_vec8_kernel<ST, 3><<<nb, 256, 0, st>>>(embedding, static_cast<const ST*>(baked), prob, grad_out,
                                                                                      grad_embedding, static_cast<ST*>(grad_baked), P, gpb, groups))); }
        else { DISPATCH_S(baked_dtype, (embed_prob_bwd_vec8_kernel<ST, 2><<<nb, 256, 0, st>>>(embedding, static_cast<const ST*>(baked), prob, grad_out,
                                                                                      grad_embedding, static_cast<ST*>(grad_baked), P, gpb, groups))); }
    } else {
        const unsigned nb = (unsigned)((P.total + 255) / 256);
        DISPATCH_S(baked_dtype, (embed_prob_bwd_kernel<ST><<<nb, 256, 0, st>>>(embedding, static_cast<const ST*>(baked), prob,
                                                                              grad_out, grad_embedding,
                                                                              static_cast<ST*>(grad_baked), P)));
    }
    SKB_LAUNCH_CHECK("embed_prob_bwd_kernel");
    return SKB_OK;
}

template <typename VT>
static int launch_vec_prob(const void* vec, const void* baked, int baked_dtype, float* prob, const float* go, void* gvec,
                           const ProbParams& P, cudaStream_t st) {
    const VT* v = static_cast<const VT*>(vec);
    VT* gv = static_cast<VT*>(gvec);
    if (vec8_ok(P.inner, {vec, baked, prob, go, gvec})) {
        const long long gpb = P.inner / 8, groups = (P.total / P.inner) * gpb;
        const unsigned nb = (unsigned)((groups + 255) / 256);
#define VP8(BWD_, NC_) DISPATCH_S(baked_dtype, (vec_prob_vec8_kernel<VT, ST, BWD_, NC_><<<nb, 256, 0, st>>>(v, static_cast<const ST*>(baked), prob, go, gv, P, gpb, groups)))
        if (go) { if (P.C == 3) { VP8(true, 3); } else { VP8(true, 2); } }
        else { if (P.C == 3) { VP8(false, 3); } else { VP8(false, 2); } }
#undef VP8
        return SKB_OK;
    }
    const unsigned nb = (unsigned)((P.total + 255) / 256);
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
// bake_skeleton: nearest skeleton point per voxel (min-reduction) + the masked 3x3x3 mean, one launch per batch
// ------------------------------------------------------------------------------------------
// Round 1 staged the WHOLE point table (up to 160 KB) into every 256-voxel CTA: ~250 B/voxel of L2 -> shared traffic for
// 16 algorithmic, occupancy capped by shared memory, one launch + one averaging launch + a blocking status read per
// sample (profiles/r01_rows.json: 1.73 ms for C4, 2 % of HBM).  Now:
//   * one launch covers the whole batch (grid.y = sample);
//   * a CTA owns an 8 x 8 x TZ tile of one sample plus, when the mean is fused, a one-voxel halo around it;
//   * it looks up the object ids that actually occur in that region (usually 0-3), and stages ONLY their skeletons into
//     shared memory with one TMA bulk copy per id (cp.async.bulk + mbarrier); ids that do not fit the arena are read
//     through the read-only cache instead — same arithmetic either way;
//   * the nearest point of every region voxel goes to shared memory, and the masked 27-mean of
//     average_baked_skeletons (skeleton.py:18-48: sum of the window / count of its entries > 0, zero padded, taps in
//     (dx,dy,dz) order) is taken from there, so the un-averaged field never reaches HBM.
// HBM traffic: the mask (+ ~1.7x halo re-reads that hit L2) in, 12 B/voxel out.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

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

constexpr int BAKE_T = 8;          // tile edge in x and y
constexpr int BAKE_MAXU = 24;      // distinct ids a CTA tracks in shared memory (more: read from global)
constexpr int BAKE_ARENA = 1536;   // skeleton points staged per CTA (24 KB)

struct BakeParams {
    int X, Y, Z, TZ, H;    // H = 1 when the 27-mean is fused (halo), else 0
    int fast_rows;         // one tile along z, rows are whole 16-byte chunks of mask and 4-float chunks of output
    int tiles_y, tiles_z;
    float an[3];
    const int* ids;        // every sample's sorted object ids, concatenated
    const int* id_begin;   // (B + 1) range of sample b in `ids`
    const int* offsets;    // (n_ids + 1) prefix of point counts, global over the batch
    const float* points;   // (n_points, 4) xyz + pad, 16-byte rows
    unsigned* status;      // SKB_STATUS_MISSING_ID: a mask id has no skeleton (the reference raises KeyError)
    const int* tri_block;  // NULL: CPU/torch semantics.  Else per sample the SKEL_BLOCK_SIZE of the reference's Triton launch
                           // (skeleton.py:361): the semantics of _min_skeleton_kernel, see bake_nearest_triton below
    long long V;           // voxels per sample
};

constexpr int BAKE_MISSING = INT_MIN;  // slot of a voxel whose id has no skeleton, Triton semantics only (no error there)

// What the reference's Triton kernel computes for one voxel (skeleton.py:171-251), which is NOT what its CPU path does:
//   * the anisotropy weighs the SQUARED differences (dist = sum_c (s_c - v_c)^2 * a_c), :208-212;
//   * the skeleton row is loaded SKEL_BLOCK_SIZE wide with a mask and no `other` (:204-206): lanes past the skeleton's
//     length hold zeros (Triton's PTX clears the destination of a predicated load), i.e. a phantom point at the origin
//     competes whenever the skeleton is shorter than the block — a voxel nearer to (0,0,0) than to its skeleton gets 0;
//   * ties: each coordinate is the MAXIMUM over the tied lanes, and over 0 from the lanes that are not tied (:219-221);
//   * an id without a skeleton silently uses index 0 with length 0 (:176,:187): all lanes phantom, result 0;
//   * baked and distance are stored as fp16 (:225-251); the distance is sqrt(min dist) (tl.sqrt is the approximate
//     square root: the fp16 result can differ by one ulp).
// Sums are taken in the written order without contraction; the reference's compiler may contract them into FMAs, which
// cannot matter while coordinates and anisotropy are integer-valued (the products are exact) — the pinned domain.
__device__ __forceinline__ void bake_nearest_triton(const float4* src, int cnt, int block, float fx, float fy, float fz,
                                                    const float (&an)[3], float& bx, float& by, float& bz, float& best) {
    int ties = 0;
    bx = by = bz = 0.f;
    best = INFINITY;
    for (int k = 0; k <= cnt; ++k) {
        if (k == cnt && cnt >= block) break;     // the block is full: no phantom lanes
        const float4 p = k < cnt ? src[k] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float ex = __fsub_rn(p.x, fx), ey = __fsub_rn(p.y, fy), ez = __fsub_rn(p.z, fz);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(ex, ex), an[0]), __fmul_rn(__fmul_rn(ey, ey), an[1])),
                                   __fmul_rn(__fmul_rn(ez, ez), an[2]));
        const int lanes = k < cnt ? 1 : block - cnt;
        if (d2 < best) { best = d2; bx = p.x; by = p.y; bz = p.z; ties = lanes; }
        else if (d2 == best) { bx = fmaxf(bx, p.x); by = fmaxf(by, p.y); bz = fmaxf(bz, p.z); ties += lanes; }
    }
    if (ties < block) { bx = fmaxf(bx, 0.f); by = fmaxf(by, 0.f); bz = fmaxf(bz, 0.f); }
    bx = __half2float(__float2half_rn(bx)); by = __half2float(__float2half_rn(by)); bz = __half2float(__float2half_rn(bz));
    best = __half2float(__float2half_rn(sqrtf(best)));
}

template <typename MT>
__global__ void __launch_bounds__(256) bake_tile_kernel(const MT* __restrict__ masks, float* __restrict__ baked,
                                                       float* __restrict__ dist_out, BakeParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int RX = BAKE_T + 2 * P.H, RY = BAKE_T + 2 * P.H, RZ = P.TZ + 2 * P.H, RN = RX * RY * RZ;
    float4* s_arena = reinterpret_cast<float4*>(smem_raw);                 // BAKE_ARENA points
    float* s_b = reinterpret_cast<float*>(s_arena + BAKE_ARENA);           // 4 x RN: x, y, z, distance
    int* s_slot = reinterpret_cast<int*>(s_b + 4 * RN);                    // RN
    __shared__ int s_present[BAKE_MAXU], s_present_id[BAKE_MAXU], s_lo[BAKE_MAXU], s_cnt[BAKE_MAXU], s_at[BAKE_MAXU];
    __shared__ __align__(8) ull bar;

    const int b = blockIdx.y;
    int t = blockIdx.x;
    const int tz = t % P.tiles_z; t /= P.tiles_z;
    const int ty = t % P.tiles_y, tx = t / P.tiles_y;
    const int x0 = tx * BAKE_T - P.H, y0 = ty * BAKE_T - P.H, z0 = tz * P.TZ - P.H;  // region origin (may be -1)
    const MT* mask = masks + (long long)b * P.V;
    const int id_lo = __ldg(P.id_begin + b), id_hi = __ldg(P.id_begin + b + 1);

    if (threadIdx.x < BAKE_MAXU) { s_present[threadIdx.x] = -1; s_present_id[threadIdx.x] = 0; }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // 0. most tiles of an instance mask hold no object at all (and neither does their halo).  When a region row is the
    //    whole z row and 16-byte chunks are legal, test it with vector loads and, if empty, store the zeros as vectors.
    if (P.fast_rows) {
        constexpr int PER = 16 / (int)sizeof(MT);
        const int chunks = P.Z / PER;
        int any = 0;
        for (int i = threadIdx.x; i < RX * RY * chunks; i += 256) {
            const int ch = i % chunks, row = i / chunks, ry = row % RY, rx = row / RY;
            const int gx = x0 + rx, gy = y0 + ry;
            if (gx >= 0 && gx < P.X && gy >= 0 && gy < P.Y) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(mask + ((long long)gx * P.Y + gy) * P.Z) + ch);
                any |= (q.x | q.y | q.z | q.w) != 0u;
            }
        }
        if (!__syncthreads_or(any)) {
            const int c4 = P.Z / 4;
            float* out0 = baked + (long long)b * 3 * P.V;
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = threadIdx.x; i < BAKE_T * BAKE_T * c4; i += 256) {
                const int ch = i % c4, row = i / c4, ly = row % BAKE_T, lx = row / BAKE_T;
                const int gx = tx * BAKE_T + lx, gy = ty * BAKE_T + ly;
                if (gx >= P.X || gy >= P.Y) continue;
                const long long g = ((long long)gx * P.Y + gy) * P.Z + 4 * ch;
                *reinterpret_cast<float4*>(out0 + g) = zero;
                *reinterpret_cast<float4*>(out0 + g + P.V) = zero;
                *reinterpret_cast<float4*>(out0 + g + 2 * P.V) = zero;
                if (dist_out) *reinterpret_cast<float4*>(dist_out + (long long)b * P.V + g) = zero;
            }
            return;
        }
    }

    // A. region voxels -> slot of their object in the CTA's id list (-1 background, <= -2: id index -2-v, list full)
    int any_fg = 0;
    constexpr int BAKE_A_BATCH = 4;  // mask loads in flight per thread (one per iteration left ~9 dependent round trips)
    for (int i0 = threadIdx.x; i0 < RN; i0 += 256 * BAKE_A_BATCH) {
        int ids_[BAKE_A_BATCH];
#pragma unroll
        for (int u = 0; u < BAKE_A_BATCH; ++u) {
            const int i = i0 + 256 * u;
            ids_[u] = 0;
            if (i < RN) {
                const int rz = i % RZ, q = i / RZ, ry = q % RY, rx = q / RY;
                const int gx = x0 + rx, gy = y0 + ry, gz = z0 + rz;
                if (gx >= 0 && gx < P.X && gy >= 0 && gy < P.Y && gz >= 0 && gz < P.Z)
                    ids_[u] = (int)mask[((long long)gx * P.Y + gy) * P.Z + gz];
            }
        }
#pragma unroll
        for (int u = 0; u < BAKE_A_BATCH; ++u) {
            const int i = i0 + 256 * u;
            if (i >= RN) break;
            const int id = ids_[u];
            int slot = -1;
            if (id != 0) {
                // a tile sees one to three objects: most voxels find their id among the ones the CTA has already looked
                // up (shared memory) and skip the binary search through the id table (five dependent global loads)
                for (int k = 0; k < BAKE_MAXU; ++k) {
                    const int pid = *reinterpret_cast<volatile int*>(&s_present_id[k]);
                    if (pid == id) { slot = k; break; }
                    if (pid == 0) break;
                }
                if (slot >= 0) {
                    any_fg = 1;
                } else {
                    int a = id_lo, e = id_hi - 1, found = -1;
                    while (a <= e) {
                        const int m = (a + e) >> 1, v = __ldg(P.ids + m);
                        if (v == id) { found = m; break; }
                        if (v < id) a = m + 1; else e = m - 1;
                    }
                    if (found < 0) {
                        if (P.tri_block) { slot = BAKE_MISSING; any_fg = 1; }
                        else atomicOr(P.status, SKB_STATUS_MISSING_ID);
                    } else {
                        any_fg = 1;
                        slot = -2 - found;
                        for (int k = 0; k < BAKE_MAXU; ++k) {
                            int cur = *reinterpret_cast<volatile int*>(&s_present[k]);
                            if (cur == -1) cur = atomicCAS(&s_present[k], -1, found);
                            if (cur == -1 || cur == found) {
                                slot = k;
                                s_present_id[k] = id;  // published after the entry: a reader that misses it just searches
                                break;
                            }
                        }
                    }
                }
            }
            s_slot[i] = slot;
        }
    }
    // most tiles of an instance mask hold no object at all (and neither does their halo): their output is zero
    if (!__syncthreads_or(any_fg)) {
        const int TN0 = BAKE_T * BAKE_T * P.TZ;
        float* out0 = baked + (long long)b * 3 * P.V;
        for (int i = threadIdx.x; i < TN0; i += 256) {
            const int lz = i % P.TZ, q = i / P.TZ, ly = q % BAKE_T, lx = q / BAKE_T;
            const int gx = tx * BAKE_T + lx, gy = ty * BAKE_T + ly, gz = tz * P.TZ + lz;
            if (gx >= P.X || gy >= P.Y || gz >= P.Z) continue;
            const long long g = ((long long)gx * P.Y + gy) * P.Z + gz;
            out0[g] = 0.f; out0[g + P.V] = 0.f; out0[g + 2 * P.V] = 0.f;
            if (dist_out) dist_out[(long long)b * P.V + g] = 0.f;
        }
        return;
    }

    // B. the present skeletons are laid out in the arena (one lane per id reads its range, thread 0 assigns the
    //    offsets and announces the byte count), then one TMA bulk copy per skeleton
    if (threadIdx.x < BAKE_MAXU && s_present[threadIdx.x] >= 0) {
        const int f = s_present[threadIdx.x];
        const int lo = __ldg(P.offsets + f);
        s_lo[threadIdx.x] = lo;
        s_cnt[threadIdx.x] = __ldg(P.offsets + f + 1) - lo;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int at = 0;
        unsigned bytes = 0;
        for (int k = 0; k < BAKE_MAXU && s_present[k] >= 0; ++k) {
            const int cnt = s_cnt[k];
            if (cnt > 0 && at + cnt <= BAKE_ARENA) { s_at[k] = at; at += cnt; bytes += (unsigned)cnt * 16u; }
            else s_at[k] = -1;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    }
    __syncthreads();          // s_at visible, the byte count announced
    if (threadIdx.x < BAKE_MAXU && s_present[threadIdx.x] >= 0 && s_at[threadIdx.x] >= 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(s_arena + s_at[threadIdx.x])),
                     "l"(P.points + 4ll * s_lo[threadIdx.x]), "r"((unsigned)s_cnt[threadIdx.x] * 16u), "r"(smem_u32(&bar))
                     : "memory");
    mbar_wait(&bar, 0);       // the staged skeletons have landed

    // C. nearest point of every region voxel (first minimum of the sqrt'ed distances, like cdist + argmin)
    const float4* gpts = reinterpret_cast<const float4*>(P.points);
    for (int i = threadIdx.x; i < RN; i += 256) {
        const int slot = s_slot[i];
        float bx = 0.f, by = 0.f, bz = 0.f, best = INFINITY;
        if (slot != -1) {
            const int rz = i % RZ, q = i / RZ, ry = q % RY, rx = q / RY;
            const float ax = __fmul_rn(P.an[0], (float)(x0 + rx)), ay = __fmul_rn(P.an[1], (float)(y0 + ry)),
                        az = __fmul_rn(P.an[2], (float)(z0 + rz));
            int lo = 0, cnt = 0;
            const float4* src = gpts;
            if (slot >= 0) {
                lo = s_lo[slot]; cnt = s_cnt[slot];
                src = s_at[slot] >= 0 ? s_arena + s_at[slot] : gpts + lo;
            } else if (slot != BAKE_MISSING) {
                const int f = -2 - slot;
                lo = __ldg(P.offsets + f); cnt = __ldg(P.offsets + f + 1) - lo;
                src = gpts + lo;
            }
            if (P.tri_block) {
                const int block = __ldg(P.tri_block + b);
                if (block > 0)
                    bake_nearest_triton(src, cnt, block, (float)(x0 + rx), (float)(y0 + ry), (float)(z0 + rz), P.an, bx, by, bz, best);
                cnt = 0;
                if (block <= 0) best = 0.f;
            }
            for (int k = 0; k < cnt; ++k) {
                const float4 p = src[k];
                const float dx = __fsub_rn(__fmul_rn(p.x, P.an[0]), ax);
                const float dy = __fsub_rn(__fmul_rn(p.y, P.an[1]), ay);
                const float dz = __fsub_rn(__fmul_rn(p.z, P.an[2]), az);
                const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                const float d = sqrtf(d2);
                if (d < best) { best = d; bx = p.x; by = p.y; bz = p.z; }
            }
            if (cnt <= 0 && !P.tri_block) best = 0.f;
        } else {
            best = 0.f;
        }
        s_b[i] = bx; s_b[RN + i] = by; s_b[2 * RN + i] = bz; s_b[3 * RN + i] = best;
    }
    __syncthreads();

    // D. outputs of the tile: the nearest point itself, or its masked 27-mean
    const int TN = BAKE_T * BAKE_T * P.TZ;
    float* out = baked + (long long)b * 3 * P.V;
    for (int i = threadIdx.x; i < TN; i += 256) {
        const int lz = i % P.TZ, q = i / P.TZ, ly = q % BAKE_T, lx = q / BAKE_T;
        const int gx = tx * BAKE_T + lx, gy = ty * BAKE_T + ly, gz = tz * P.TZ + lz;
        if (gx >= P.X || gy >= P.Y || gz >= P.Z) continue;
        const long long g = ((long long)gx * P.Y + gy) * P.Z + gz;
        const int c = ((lx + P.H) * RY + (ly + P.H)) * RZ + (lz + P.H);
        if (P.H == 0) {
            out[g] = s_b[c]; out[g + P.V] = s_b[RN + c]; out[g + 2 * P.V] = s_b[2 * RN + c];
        } else {
            // a window without any object voxel averages to zero (slots: -1 = background)
            int any = 0;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dz = -1; dz <= 1; ++dz) any |= (s_slot[c + (dx * RY + dy) * RZ + dz] != -1);
            if (!any) {
                out[g] = 0.f; out[g + P.V] = 0.f; out[g + 2 * P.V] = 0.f;
                if (dist_out) dist_out[(long long)b * P.V + g] = 0.f;
                continue;
            }
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float* sb = s_b + ch * RN;
                float sum = 0.f, cnt = 0.f;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                        for (int dz = -1; dz <= 1; ++dz) {
                            const float v = sb[c + (dx * RY + dy) * RZ + dz];  // 0 outside the volume, as zero padding
                            sum = __fadd_rn(sum, v);
                            cnt += v > 0.f ? 1.f : 0.f;
                        }
                out[g + ch * P.V] = __fdiv_rn(sum, cnt == 0.f ? 1.f : cnt);
            }
        }
        if (dist_out) dist_out[(long long)b * P.V + g] = s_b[3 * RN + c];
    }
}

extern "C" int skb_bake_skeletons(const void* masks, int mask_dtype, int64_t B, int64_t X, int64_t Y, int64_t Z,
                                  const int32_t* ids, const int32_t* id_begin, const int32_t* offsets, int n_ids,
                                  const float* points_xyzw, int n_points, const float anisotropy[3], int average,
                                  const int32_t* triton_block, float* baked, float* distance, uint32_t* status, void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_bake_skeletons");
    if (rc) return rc;
    SKB_REQUIRE(masks && baked && status && anisotropy && id_begin && B >= 1 && B <= 65535 && n_ids >= 0 && n_points >= 0,
                "skb_bake_skeletons: bad argument");
    SKB_REQUIRE(n_ids == 0 || (ids && offsets), "skb_bake_skeletons: NULL id tables");
    SKB_REQUIRE(n_points == 0 || (points_xyzw && skb_aligned16(points_xyzw)), "skb_bake_skeletons: point table must be 16-byte aligned");
    SKB_REQUIRE(mask_dtype == SKB_I32 || mask_dtype == SKB_I16 || mask_dtype == SKB_U8, "skb_bake_skeletons: mask dtype must be u8, i16 or i32");
    BakeParams P;
    P.X = (int)X; P.Y = (int)Y; P.Z = (int)Z;
    P.H = average ? 1 : 0;
    P.TZ = Z < 32 ? (int)Z : 32;
    const int tiles_x = (int)((X + BAKE_T - 1) / BAKE_T);
    P.tiles_y = (int)((Y + BAKE_T - 1) / BAKE_T);
    P.tiles_z = (int)((Z + P.TZ - 1) / P.TZ);
    P.an[0] = anisotropy[0]; P.an[1] = anisotropy[1]; P.an[2] = anisotropy[2];
    P.ids = ids; P.id_begin = id_begin; P.offsets = offsets; P.points = points_xyzw; P.status = status;
    P.tri_block = triton_block;
    P.V = X * Y * Z;
    const long long tiles = (long long)tiles_x * P.tiles_y * P.tiles_z;
    SKB_REQUIRE(tiles < (1LL << 31), "skb_bake_skeletons: too many tiles");
    const int msize = mask_dtype == SKB_I32 ? 4 : (mask_dtype == SKB_I16 ? 2 : 1);
    P.fast_rows = (P.tiles_z == 1 && (Z * msize) % 16 == 0 && Z % 4 == 0 && skb_aligned16(masks) && skb_aligned16(baked) &&
                   (!distance || skb_aligned16(distance))) ? 1 : 0;
    const int RN = (BAKE_T + 2 * P.H) * (BAKE_T + 2 * P.H) * (P.TZ + 2 * P.H);
    const size_t smem = (size_t)BAKE_ARENA * 16 + (size_t)RN * 5 * 4;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid((unsigned)tiles, (unsigned)B);
#define BAKE_LAUNCH(MT)                                                                                               \
    do {                                                                                                              \
        SKB_RAISE_SMEM_ONCE(bake_tile_kernel<MT>, 200 * 1024);                                                        \
        bake_tile_kernel<MT><<<grid, 256, smem, st>>>(static_cast<const MT*>(masks), baked, distance, P);             \
    } while (0)
    if (mask_dtype == SKB_I32) BAKE_LAUNCH(int32_t);
    else if (mask_dtype == SKB_I16) BAKE_LAUNCH(int16_t);
    else BAKE_LAUNCH(uint8_t);
    SKB_LAUNCH_CHECK("bake_tile_kernel");
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
