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

// Six consecutive z inputs (z0-1 .. z0+4) of one row, zero outside the volume.  When the row is 16-byte aligned and the
// four middle elements exist, they come in as one 16-byte load.
__device__ __forceinline__ void load_row6(const float* __restrict__ row, int z0, int Z, bool aligned4, float (&v)[6]) {
    if (aligned4 && z0 + 4 <= Z) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(row + z0));
        v[1] = m.x; v[2] = m.y; v[3] = m.z; v[4] = m.w;
        v[0] = z0 > 0 ? __ldg(row + z0 - 1) : 0.f;
        v[5] = z0 + 4 < Z ? __ldg(row + z0 + 4) : 0.f;
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int zz = z0 - 1 + k;
            v[k] = (zz >= 0 && zz < Z) ? __ldg(row + zz) : 0.f;
        }
    }
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
    const bool aligned4 = (Z % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15u) == 0);

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
                load_row6(base + ((long long)xx * Y + yy) * Z, z0, Z, aligned4, v);
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
    if (aligned4 && z0 + 4 <= Z && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        *reinterpret_cast<float4*>(orow + z0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (z0 + j < Z) orow[z0 + j] = acc[j];
    }
}

// Round 1 issued 27 scalar loads per voxel.  Now one thread produces 4 consecutive z outputs from 9 rows of 6 inputs
// (54 loads -> 13.5 per voxel, mostly 16-byte ones).  The reference sums the 27 conv channels in tap order
// (dx, dy, dz ascending) with zeros at the border: the same order is kept per output.
__global__ void __launch_bounds__(256) masked_mean27_kernel(const float* __restrict__ in, float* __restrict__ out, int X,
                                                           int Y, int Z, long long n_groups, int Z4) {
    const long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_groups) return;
    const int zg = (int)(gi % Z4);
    long long r = gi / Z4;
    const int y = (int)(r % Y);
    r /= Y;
    const int x = (int)(r % X);
    const long long vol = r / X;
    const int z0 = zg * 4;
    const float* base = in + vol * (long long)X * Y * Z;
    const bool aligned4 = (Z % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15u) == 0);
    float sum[4] = {0.f, 0.f, 0.f, 0.f}, cnt[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int xx = x + dx, yy = y + dy;
            float v[6];
            if (xx < 0 || xx >= X || yy < 0 || yy >= Y) {
#pragma unroll
                for (int k = 0; k < 6; ++k) v[k] = 0.f;
            } else {
                load_row6(base + ((long long)xx * Y + yy) * Z, z0, Z, aligned4, v);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int dz = 0; dz < 3; ++dz) {
                    sum[j] = __fadd_rn(sum[j], v[j + dz]);
                    cnt[j] += v[j + dz] > 0.f ? 1.f : 0.f;
                }
            }
        }
    }
    float* orow = out + vol * (long long)X * Y * Z + ((long long)x * Y + y) * Z;
    if (aligned4 && z0 + 4 <= Z && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        float4 o;
        o.x = __fdiv_rn(sum[0], cnt[0] == 0.f ? 1.f : cnt[0]);
        o.y = __fdiv_rn(sum[1], cnt[1] == 0.f ? 1.f : cnt[1]);
        o.z = __fdiv_rn(sum[2], cnt[2] == 0.f ? 1.f : cnt[2]);
        o.w = __fdiv_rn(sum[3], cnt[3] == 0.f ? 1.f : cnt[3]);
        *reinterpret_cast<float4*>(orow + z0) = o;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (z0 + j < Z) orow[z0 + j] = __fdiv_rn(sum[j], cnt[j] == 0.f ? 1.f : cnt[j]);
    }
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

// Round 1 took the 7 x 7 x 3 box max by brute force: 147 taps x 2 loads per output voxel (84 us per 300x300x20 tile).
// A box max is separable: a CTA owns a 16 x 16 x bz block of interior outputs, stages the masked skeleton values of the
// block plus its (3,3,1) apron in shared memory once (-inf outside the tile: it never wins), and reduces along z (3 taps),
// y (7) and x (7) — the z pass in registers while staging, y and x in shared memory — 2 global loads per staged voxel
// instead of 294 per output.
constexpr int EPI_BX = 16, EPI_BY = 16, EPI_BZ = 16, EPI_RX = EPI_BX + 6, EPI_RY = EPI_BY + 6;

template <typename T>
__global__ void __launch_bounds__(256) tile_epilogue_kernel(EpiParams P, int bz, int nby, int nbz) {
    extern __shared__ float epi_smem[];
    float* B = epi_smem;                          // [RX][RY][bz]  masked skeleton after the z pass
    float* A = epi_smem + EPI_RX * EPI_RY * bz;   // [RX][BY][bz]  after the y pass
    int t = blockIdx.x;
    const int kz = t % nbz; t /= nbz;
    const int ky = t % nby, kx = t / nby;
    const int ix = P.tx - 2 * P.mx, iy = P.ty - 2 * P.my, iz = P.tz - 2 * P.mz;
    // tile-local coordinates of the block's first output voxel
    const int lx0 = P.mx + kx * EPI_BX, ly0 = P.my + ky * EPI_BY, lz0 = P.mz + kz * bz;
    const long long plane = (long long)P.tx * P.ty * P.tz;
    const void* prob = static_cast<const char*>(P.unet) + (size_t)(P.C - 1) * plane * sizeof(T);
    const void* skel = static_cast<const char*>(P.unet) + (size_t)(P.C - 2) * plane * sizeof(T);

    // stage + z pass: one thread per (x, y) row of the block + apron.  It reads the row's bz + 2 values of
    // (skel.float() * (prob > thr)) (eval.py:146-150; -inf outside the tile: never wins) — all loads of a row are
    // independent and issued together — and stores the 3-tap z maxima.
    for (int row = threadIdx.x; row < EPI_RX * EPI_RY; row += 256) {
        const int y = row % EPI_RY, x = row / EPI_RY;
        const int xx = lx0 - 3 + x, yy = ly0 - 3 + y;
        float v[EPI_BZ + 2];
        const bool inside = xx >= 0 && xx < P.tx && yy >= 0 && yy < P.ty;
        const long long at0 = ((long long)xx * P.ty + yy) * P.tz;
#pragma unroll
        for (int k = 0; k < EPI_BZ + 2; ++k) {
            const int zz = lz0 - 1 + k;
            float s_ = 0.f, p_ = 0.f;
            const bool ok = inside && k < bz + 2 && zz >= 0 && zz < P.tz;
            if (ok) { s_ = epi_load<T>(skel, at0 + zz); p_ = epi_load<T>(prob, at0 + zz); }
            v[k] = ok ? s_ * (p_ > P.thr ? 1.f : 0.f) : -INFINITY;
        }
#pragma unroll
        for (int k = 0; k < EPI_BZ; ++k)
            if (k < bz) B[row * bz + k] = fmaxf(fmaxf(v[k], v[k + 1]), v[k + 2]);
    }
    __syncthreads();
    const int n_y = EPI_RX * EPI_BY * bz;
    for (int i = threadIdx.x; i < n_y; i += 256) {
        const int z = i % bz, q = i / bz, y = q % EPI_BY, x = q / EPI_BY;
        const float* b0 = B + (x * EPI_RY + y) * bz + z;
        float m = b0[0];
#pragma unroll
        for (int d = 1; d <= 6; ++d) m = fmaxf(m, b0[d * bz]);
        A[i] = m;
    }
    __syncthreads();
    const int n_out = EPI_BX * EPI_BY * bz;
    const long long V = (long long)P.X * P.Y * P.Z;
    // Outputs in batches of EPI_BATCH per thread: the batch's global loads (probability + three vector channels) are all
    // issued before anything is stored — one output per iteration left this loop waiting on a load round trip per output
    // (10 dependent round trips per thread, about half of the kernel's time).
    constexpr int EPI_BATCH = 5;
    for (int i0 = threadIdx.x; i0 < n_out; i0 += 256 * EPI_BATCH) {
        float pv[EPI_BATCH], vv[EPI_BATCH][3], mm[EPI_BATCH];
        long long gat[EPI_BATCH];
        bool ok[EPI_BATCH];
#pragma unroll
        for (int u = 0; u < EPI_BATCH; ++u) {
            const int i = i0 + 256 * u;
            const int z = i % bz, q = i / bz, y = q % EPI_BY, x = q / EPI_BY;
            const int lx = lx0 + x, ly = ly0 + y, lz = lz0 + z;
            ok[u] = i < n_out && lx < P.mx + ix && ly < P.my + iy && lz < P.mz + iz;
            pv[u] = 0.f; vv[u][0] = vv[u][1] = vv[u][2] = 0.f; mm[u] = 0.f; gat[u] = 0;
            if (ok[u]) {
                const long long at = ((long long)lx * P.ty + ly) * P.tz + lz;
                pv[u] = epi_load<T>(prob, at);
#pragma unroll
                for (int c = 0; c < 3; ++c) vv[u][c] = epi_load<T>(P.unet, c * plane + at);
                const float* c0 = A + (x * EPI_BY + y) * bz + z;
                float m = c0[0];
#pragma unroll
                for (int d = 1; d <= 6; ++d) m = fmaxf(m, c0[d * EPI_BY * bz]);
                // zero joins the max only where the window of one of the three zero-padded dilations leaves the tile
                if (lx <= 2 || lx >= P.tx - 3 || ly <= 2 || ly >= P.ty - 3 || lz == 0 || lz == P.tz - 1) m = fmaxf(m, 0.f);
                mm[u] = m;
                gat[u] = ((long long)(P.ox + lx) * P.Y + (P.oy + ly)) * P.Z + (P.oz + lz);
            }
        }
#pragma unroll
        for (int u = 0; u < EPI_BATCH; ++u) {
            if (!ok[u]) continue;
            P.skel[gat[u]] = mm[u] > P.thr ? 1 : 0;
            // vectors: v * (prob > thr) computed in the network dtype, then .half()  (eval.py:149,175)
            const float keep = pv[u] > P.thr ? 1.f : 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v = vv[u][c] * keep;
                v = skb_to_float<T>(skb_from_float<T>(v));  // the product is rounded to the network dtype first
                P.vectors[c * V + gat[u]] = __float2half_rn(v);
            }
        }
    }
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
    const int Z4 = (int)((Z + 3) / 4);
    const long long groups = n_volumes * X * Y * Z4;
    masked_mean27_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out, (int)X, (int)Y, (int)Z, groups, Z4);
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
    const int ix = P.tx - 2 * P.mx, iy = P.ty - 2 * P.my, iz = P.tz - 2 * P.mz;
    const int bz = iz < EPI_BZ ? iz : EPI_BZ;
    const int nbx = (ix + EPI_BX - 1) / EPI_BX, nby = (iy + EPI_BY - 1) / EPI_BY, nbz = (iz + bz - 1) / bz;
    const long long blocks = (long long)nbx * nby * nbz;
    SKB_REQUIRE(blocks < (1LL << 31), "skb_tile_epilogue: tile too large");
    const int smem = (int)sizeof(float) * (EPI_RX * EPI_RY + EPI_RX * EPI_BY) * bz;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define EPI_LAUNCH(T)                                                                                          \
    do {                                                                                                       \
        SKB_RAISE_SMEM_ONCE(tile_epilogue_kernel<T>, 96 * 1024);                                               \
        tile_epilogue_kernel<T><<<(unsigned)blocks, 256, smem, st>>>(P, bz, nby, nbz);                         \
    } while (0)
    if (in_dtype == SKB_F32) EPI_LAUNCH(float);
    else if (in_dtype == SKB_F16) EPI_LAUNCH(__half);
    else EPI_LAUNCH(__nv_bfloat16);
#undef EPI_LAUNCH
    SKB_LAUNCH_CHECK("tile_epilogue_kernel");
    return SKB_OK;
}
