// Stencil kernels around the assembly path (kernel (d) of the north star).
//
//   skb_stencil3      3x3x3 max / 3x3x1 max / 3x3x3 min with ZERO padding — the arithmetic of
//                     skoots/lib/morphology.py:130-199, which the reference performs as a 27-channel
//                     one-hot conv3d (27x the input in HBM) followed by a channel reduction.
//   skb_masked_mean27 per-channel sum(window)/max(1,count(window>0)), zero padded —
//                     skoots/lib/skeleton.py:18-48 (average_baked_skeletons).
//   skb_tile_epilogue skoots/lib/eval.py:145-176 fused: prob>thr masks vectors and skeleton,
//                     one 3x3x3 and two 3x3x1 dilations, skeleton>thr -> u8, vectors -> fp16,
//                     interior of the tile written into the whole-volume arrays.
//
// One thread produces 4 consecutive z outputs: per (dx,dy) row it loads the 6 inputs it needs
// (one 16-byte load + 2 scalars when the row is aligned) and reduces along z in registers, then
// across the 9 (or 3x3, or 1) rows.  HBM sees each input once (8 B/voxel fp32 in+out); the 9x
// re-reads are L1/L2 hits.
#include "skb_common.cuh"

enum { OP_MAX333 = 0, OP_MAX331 = 1, OP_MIN333 = 2 };

template <int OP>
__device__ __forceinline__ float red(float a, float b) {
    return OP == OP_MIN333 ? fminf(a, b) : fmaxf(a, b);
}

// torch.max/min propagate NaN; fmaxf/fminf do not.  The reference's conv3d already turns NaN into
// NaN for the whole window (0*NaN), so NaN inputs are outside the parity contract.
template <int OP>
__global__ void __launch_bounds__(256) stencil3_kernel(const float* __restrict__ in, float* __restrict__ out, int X,
                                                      int Y, int Z, long long n_groups, int Z4) {
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_groups) return;
    const int zg = (int)(gi % Z4);
    long long r = gi / Z4;
    const int y = (int)(r % Y);
    r /= Y;
    const int x = (int)(r % X);
    const long long vol = r / X;  // flattened (b,c)
    const int z0 = zg * 4;
    const float* base = in + vol * (long long)X * Y * Z;
    constexpr int RZ = OP == OP_MAX331 ? 0 : 1;

    float acc[4];
    bool first = true;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int xx = x + dx, yy = y + dy;
            float v[6];  // inputs z0-1 .. z0+4, zero outside the volume
            if (xx < 0 || xx >= X || yy < 0 || yy >= Y) {
#pragma unroll
                for (int k = 0; k < 6; ++k) v[k] = 0.f;
            } else {
                const float* row = base + ((long long)xx * Y + yy) * Z;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int zz = z0 - 1 + k;
                    v[k] = (zz >= 0 && zz < Z) ? __ldg(row + zz) : 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float m = v[j + 1];
                if (RZ) m = red<OP>(red<OP>(v[j], m), v[j + 2]);
                acc[j] = first ? m : red<OP>(acc[j], m);
            }
            first = false;
        }
    }
    float* orow = out + vol * (long long)X * Y * Z + ((long long)x * Y + y) * Z;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (z0 + j < Z) orow[z0 + j] = acc[j];
}

__global__ void __launch_bounds__(256) masked_mean27_kernel(const float* __restrict__ in, float* __restrict__ out, int X,
                                                           int Y, int Z, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int z = (int)(i % Z);
    long long r = i / Z;
    const int y = (int)(r % Y);
    r /= Y;
    const int x = (int)(r % X);
    const float* base = in + (r / X) * (long long)X * Y * Z;
    // the reference sums the 27 conv channels in tap order (dx,dy,dz ascending) with zeros at the border
    float sum = 0.f, cnt = 0.f;
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dz = -1; dz <= 1; ++dz) {
                const int xx = x + dx, yy = y + dy, zz = z + dz;
                float v = 0.f;
                if (xx >= 0 && xx < X && yy >= 0 && yy < Y && zz >= 0 && zz < Z)
                    v = __ldg(base + ((long long)xx * Y + yy) * Z + zz);
                sum = __fadd_rn(sum, v);
                cnt += v > 0.f ? 1.f : 0.f;
            }
    out[i] = __fdiv_rn(sum, cnt == 0.f ? 1.f : cnt);
}

// ---- tile epilogue -------------------------------------------------------------------------------
struct EpiParams {
    const void* unet;     // (C, x, y, z) of one tile (batch 1)
    int in_dtype;         // SKB_F32 | SKB_F16 | SKB_BF16
    int C, tx, ty, tz;    // tile dims
    int ox, oy, oz;       // tile origin in the volume
    int mx, my, mz;       // overlap margins trimmed from each side
    float thr;
    __half* vectors;      // (3, X, Y, Z)
    unsigned char* skel;  // (X, Y, Z)
    int X, Y, Z;
};

template <typename T>
__device__ __forceinline__ float epi_load(const void* base, long long idx) {
    return skb_to_float<T>(static_cast<const T*>(base)[idx]);
}

template <typename T>
__global__ void __launch_bounds__(256) tile_epilogue_kernel(EpiParams P, long long n_interior) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_interior) return;
    const int iz = P.tz - 2 * P.mz, iy = P.ty - 2 * P.my;
    const int lz = (int)(i % iz) + P.mz;
    long long r = i / iz;
    const int ly = (int)(r % iy) + P.my;
    const int lx = (int)(r / iy) + P.mx;
    const long long plane = (long long)P.tx * P.ty * P.tz;
    const long long at = ((long long)lx * P.ty + ly) * P.tz + lz;
    const void* prob = static_cast<const char*>(P.unet) + (size_t)(P.C - 1) * plane * sizeof(T);
    const void* skel = static_cast<const char*>(P.unet) + (size_t)(P.C - 2) * plane * sizeof(T);

    // vectors: v * (prob > thr) computed in the network dtype, then .half()  (eval.py:149,175)
    const float keep = epi_load<T>(prob, at) > P.thr ? 1.f : 0.f;
    const long long gx = P.ox + lx, gy = P.oy + ly, gz = P.oz + lz;
    const long long gat = (gx * P.Y + gy) * P.Z + gz, V = (long long)P.X * P.Y * P.Z;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = epi_load<T>(P.unet, c * plane + at) * keep;
        v = skb_to_float<T>(skb_from_float<T>(v));  // the product is rounded to the network dtype first
        P.vectors[c * V + gat] = __float2half_rn(v);
    }

    // skeleton: (skel.float() * (prob>thr)) dilated 3x3x3 then 3x3x1 twice == zero-padded max over
    // |dx|,|dy| <= 3, |dz| <= 1; zero joins the max only where a stage's window leaves the tile
    float m = 0.f;
    bool have = (lx <= 2 || lx >= P.tx - 3 || ly <= 2 || ly >= P.ty - 3 || lz == 0 || lz == P.tz - 1);
    for (int dx = -3; dx <= 3; ++dx) {
        const int xx = lx + dx;
        if (xx < 0 || xx >= P.tx) continue;
        for (int dy = -3; dy <= 3; ++dy) {
            const int yy = ly + dy;
            if (yy < 0 || yy >= P.ty) continue;
            for (int dz = -1; dz <= 1; ++dz) {
                const int zz = lz + dz;
                if (zz < 0 || zz >= P.tz) continue;
                const long long q = ((long long)xx * P.ty + yy) * P.tz + zz;
                const float s = epi_load<T>(skel, q) * (epi_load<T>(prob, q) > P.thr ? 1.f : 0.f);
                m = have ? fmaxf(m, s) : s;
                have = true;
            }
        }
    }
    P.skel[gat] = m > P.thr ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
extern "C" int skb_stencil3(const float* in, float* out, int64_t n_volumes, int64_t X, int64_t Y, int64_t Z, int op,
                            void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_stencil3");
    if (rc) return rc;
    SKB_REQUIRE(in && out && in != out && n_volumes >= 1, "skb_stencil3: bad argument (in-place is not supported)");
    SKB_REQUIRE(op >= 0 && op <= 2, "skb_stencil3: op must be 0 (max 3x3x3), 1 (max 3x3x1) or 2 (min 3x3x3)");
    const int Z4 = (int)((Z + 3) / 4);
    const long long groups = n_volumes * X * Y * Z4;
    const unsigned nb = (unsigned)((groups + 255) / 256);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (op == OP_MAX333) stencil3_kernel<OP_MAX333><<<nb, 256, 0, st>>>(in, out, (int)X, (int)Y, (int)Z, groups, Z4);
    else if (op == OP_MAX331) stencil3_kernel<OP_MAX331><<<nb, 256, 0, st>>>(in, out, (int)X, (int)Y, (int)Z, groups, Z4);
    else stencil3_kernel<OP_MIN333><<<nb, 256, 0, st>>>(in, out, (int)X, (int)Y, (int)Z, groups, Z4);
    SKB_LAUNCH_CHECK("stencil3_kernel");
    return SKB_OK;
}

extern "C" int skb_masked_mean27(const float* in, float* out, int64_t n_volumes, int64_t X, int64_t Y, int64_t Z,
                                 void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_masked_mean27");
    if (rc) return rc;
    SKB_REQUIRE(in && out && in != out && n_volumes >= 1, "skb_masked_mean27: bad argument");
    const long long total = n_volumes * X * Y * Z;
    masked_mean27_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out, (int)X, (int)Y, (int)Z, total);
    SKB_LAUNCH_CHECK("masked_mean27_kernel");
    return SKB_OK;
}

extern "C" int skb_tile_epilogue(const void* unet, int in_dtype, int C, const int32_t tile[3], const int32_t origin[3],
                                 const int32_t overlap[3], float threshold, void* vectors_f16, uint8_t* skeleton_u8,
                                 int64_t X, int64_t Y, int64_t Z, void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_tile_epilogue");
    if (rc) return rc;
    SKB_REQUIRE(unet && tile && origin && overlap && vectors_f16 && skeleton_u8, "skb_tile_epilogue: NULL pointer");
    SKB_REQUIRE(C >= 5, "skb_tile_epilogue: the network output needs >= 5 channels (3 vector, skeleton, probability)");
    SKB_REQUIRE(in_dtype == SKB_F32 || in_dtype == SKB_F16 || in_dtype == SKB_BF16, "skb_tile_epilogue: dtype");
    const int64_t dims[3] = {X, Y, Z};
    for (int a = 0; a < 3; ++a) {
        SKB_REQUIRE(tile[a] > 0 && overlap[a] >= 0 && tile[a] - 2 * overlap[a] > 0, "skb_tile_epilogue: tile/overlap");
        SKB_REQUIRE(origin[a] >= 0 && origin[a] + tile[a] <= dims[a], "skb_tile_epilogue: tile leaves the volume");
    }
    EpiParams P;
    P.unet = unet; P.in_dtype = in_dtype; P.C = C;
    P.tx = tile[0]; P.ty = tile[1]; P.tz = tile[2];
    P.ox = origin[0]; P.oy = origin[1]; P.oz = origin[2];
    P.mx = overlap[0]; P.my = overlap[1]; P.mz = overlap[2];
    P.thr = threshold;
    P.vectors = static_cast<__half*>(vectors_f16); P.skel = skeleton_u8;
    P.X = (int)X; P.Y = (int)Y; P.Z = (int)Z;
    const long long n = (long long)(P.tx - 2 * P.mx) * (P.ty - 2 * P.my) * (P.tz - 2 * P.mz);
    const unsigned nb = (unsigned)((n + 255) / 256);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_dtype == SKB_F32) tile_epilogue_kernel<float><<<nb, 256, 0, st>>>(P, n);
    else if (in_dtype == SKB_F16) tile_epilogue_kernel<__half><<<nb, 256, 0, st>>>(P, n);
    else tile_epilogue_kernel<__nv_bfloat16><<<nb, 256, 0, st>>>(P, n);
    SKB_LAUNCH_CHECK("tile_epilogue_kernel");
    return SKB_OK;
}
