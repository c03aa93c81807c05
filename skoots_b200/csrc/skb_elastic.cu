// SURVEY §8 row f4 (the pinnable part): the elastic deformation of the training augmentation,
// skoots/train/merged_transform.py:75-188 (`elastic_deform`) and :43-72 (`_elastic_on_skeletons`).
//
// The reference draws a coarse random displacement field (1,3,dz,dy,dx), upsamples it to the crop with
// F.interpolate(mode="trilinear"), scales it, adds it to an identity grid built from three linspaces (a dense
// (1,X,Y,Z,3) fp32 tensor: 12 B/voxel, plus ~10 full-volume temporaries), resamples every argument with
// F.grid_sample(mode="nearest", align_corners=True), then builds a SECOND dense grid ((base - offset + 1)/2 * size) only
// to look up the new position of each skeleton point (a loop the author marks "INSANELY SLOW", :169).
//
// Here neither grid exists in HBM.  The displacement at a voxel is 24 values of the coarse field and 7 lerps, so both
// kernels evaluate it on the fly from the coarse field (a few hundred bytes, cached):
//   skb_elastic_resample   out[v] = in[nearest(base + offset)]     4 B read + 4 B written per voxel and argument
//   skb_elastic_points     p' = ((base - offset + 1)/2 * size)[p]  one thread per skeleton point
// Arithmetic: the reference's formulas restated in fp32 with one rounding per operation (linspace's two-sided
// evaluation, area_pixel_compute_source_index with align_corners = False, the w -> h -> d order of the lerps, the
// grid_sample un-normalisation and nearbyint).  ATen's own CPU and CUDA kernels do not agree with each other to the last
// bit on the trilinear upsampling, so parity for this row is stated with a tolerance: displacements within
// 1e-6, hence identical samples except where a coordinate falls within ~1e-4 of a rounding boundary.
#include "skb_common.cuh"

struct ElasticParams {
    const float* noise;  // (3, D, H, W): the coarse field; D runs along X, H along Y, W along Z (merged_transform.py:129-135)
    int D, H, W;
    int X, Y, Z;
    float mag[3];        // reversed displacement_magnitude: channel c scales grid component c (:136,:144)
};

// source index / weight of F.interpolate(mode="trilinear", align_corners=False) along one axis
__device__ __forceinline__ void lin_src(int dst, int in_size, int out_size, int& i0, int& i1, float& l0, float& l1) {
    if (in_size == out_size) { i0 = i1 = dst; l0 = 1.f; l1 = 0.f; return; }
    const float scale = __fdiv_rn((float)in_size, (float)out_size);
    float real = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
    real = fmaxf(real, 0.f);
    i0 = min((int)floorf(real), in_size - 1);
    l1 = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.f), 1.f);
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l0 = __fsub_rn(1.f, l1);
}

// torch.linspace(-1, 1, n)[i]: start + i*step below the middle, end - (n-1-i)*step above it
__device__ __forceinline__ float linspace_pm1(int i, int n) {
    if (n == 1) return -1.f;
    const float step = __fdiv_rn(2.f, (float)(n - 1));
    return i < n / 2 ? __fadd_rn(-1.f, __fmul_rn(step, (float)i)) : __fsub_rn(1.f, __fmul_rn(step, (float)(n - 1 - i)));
}

// the three displacement components at voxel (x, y, z), already scaled by the magnitudes
__device__ __forceinline__ void elastic_offset(const ElasticParams& P, int x, int y, int z, float (&off)[3]) {
    int d0, d1, h0, h1, w0, w1;
    float ld0, ld1, lh0, lh1, lw0, lw1;
    lin_src(x, P.D, P.X, d0, d1, ld0, ld1);
    lin_src(y, P.H, P.Y, h0, h1, lh0, lh1);
    lin_src(z, P.W, P.Z, w0, w1, lw0, lw1);
    const int plane = P.H * P.W, vol = P.D * plane;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* n = P.noise + c * vol;
        const float* r00 = n + d0 * plane + h0 * P.W;
        const float* r01 = n + d0 * plane + h1 * P.W;
        const float* r10 = n + d1 * plane + h0 * P.W;
        const float* r11 = n + d1 * plane + h1 * P.W;
        const float a00 = __fadd_rn(__fmul_rn(__ldg(r00 + w0), lw0), __fmul_rn(__ldg(r00 + w1), lw1));
        const float a01 = __fadd_rn(__fmul_rn(__ldg(r01 + w0), lw0), __fmul_rn(__ldg(r01 + w1), lw1));
        const float a10 = __fadd_rn(__fmul_rn(__ldg(r10 + w0), lw0), __fmul_rn(__ldg(r10 + w1), lw1));
        const float a11 = __fadd_rn(__fmul_rn(__ldg(r11 + w0), lw0), __fmul_rn(__ldg(r11 + w1), lw1));
        const float b0 = __fadd_rn(__fmul_rn(a00, lh0), __fmul_rn(a01, lh1));
        const float b1 = __fadd_rn(__fmul_rn(a10, lh0), __fmul_rn(a11, lh1));
        off[c] = __fmul_rn(__fadd_rn(__fmul_rn(b0, ld0), __fmul_rn(b1, ld1)), P.mag[c]);
    }
}

// grid component 0 pairs with the z axis, 1 with y, 2 with x (base_grid = stack(meshz, meshy, meshx), :143)
__global__ void __launch_bounds__(256) elastic_resample_kernel(ElasticParams P, const float* __restrict__ in,
                                                              float* __restrict__ out, long long n_vol) {
    const long long V = (long long)P.X * P.Y * P.Z;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const int z = (int)(i % P.Z);
    const long long q = i / P.Z;
    const int y = (int)(q % P.Y), x = (int)(q / P.Y);
    float off[3];
    elastic_offset(P, x, y, z, off);
    const float gz = __fadd_rn(linspace_pm1(z, P.Z), off[0]);
    const float gy = __fadd_rn(linspace_pm1(y, P.Y), off[1]);
    const float gx = __fadd_rn(linspace_pm1(x, P.X), off[2]);
    // grid_sampler_unnormalize, align_corners = True: ((g + 1) / 2) * (size - 1); nearest = nearbyint (half to even)
    const float fz = nearbyintf(__fmul_rn(__fdiv_rn(__fadd_rn(gz, 1.f), 2.f), (float)(P.Z - 1)));
    const float fy = nearbyintf(__fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(P.Y - 1)));
    const float fx = nearbyintf(__fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(P.X - 1)));
    const bool inside = fz >= 0.f && fz < (float)P.Z && fy >= 0.f && fy < (float)P.Y && fx >= 0.f && fx < (float)P.X;
    const long long src = inside ? ((long long)fx * P.Y + (long long)fy) * P.Z + (long long)fz : 0;
    for (long long v = 0; v < n_vol; ++v) out[v * V + i] = inside ? __ldg(in + v * V + src) : 0.f;  // padding_mode = zeros
}

// points (n,3) as int64 (the reference indexes the grid with them, so they are integer tensors; the assignment back
// truncates toward zero) or fp32 (kept as floats).  Points outside the volume are left alone (:57-66).
template <typename PT>
__global__ void __launch_bounds__(128) elastic_points_kernel(ElasticParams P, const PT* __restrict__ pts, PT* __restrict__ out,
                                                            long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PT px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
    PT ox = px, oy = py, oz = pz;
    if (px >= (PT)0 && px < (PT)P.X && py >= (PT)0 && py < (PT)P.Y && pz >= (PT)0 && pz < (PT)P.Z) {
        const int x = (int)px, y = (int)py, z = (int)pz;
        float off[3];
        elastic_offset(P, x, y, z, off);
        // grid = (base - offset).add(1).div(2).mul((z, y, x)); the point takes components [2, 1, 0]  (:162-168, :69)
        const float nz = __fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(linspace_pm1(z, P.Z), off[0]), 1.f), 2.f), (float)P.Z);
        const float ny = __fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(linspace_pm1(y, P.Y), off[1]), 1.f), 2.f), (float)P.Y);
        const float nx = __fmul_rn(__fdiv_rn(__fadd_rn(__fsub_rn(linspace_pm1(x, P.X), off[2]), 1.f), 2.f), (float)P.X);
        ox = (PT)nx; oy = (PT)ny; oz = (PT)nz;  // integer PT: truncation toward zero, like the reference's assignment
    }
    out[3 * i] = ox; out[3 * i + 1] = oy; out[3 * i + 2] = oz;
}

static int fill_elastic(ElasticParams& P, const char* who, const float* noise, int64_t D, int64_t H, int64_t W,
                        const float mag[3], int64_t X, int64_t Y, int64_t Z) {
    int rc = skb_check_volume(X, Y, Z, who);
    if (rc) return rc;
    if (!noise || !mag || D < 1 || H < 1 || W < 1 || D > 4096 || H > 4096 || W > 4096) {
        skb_set_error("%s: bad displacement field", who);
        return SKB_E_ARG;
    }
    P.noise = noise; P.D = (int)D; P.H = (int)H; P.W = (int)W;
    P.X = (int)X; P.Y = (int)Y; P.Z = (int)Z;
    P.mag[0] = mag[0]; P.mag[1] = mag[1]; P.mag[2] = mag[2];
    return SKB_OK;
}

extern "C" int skb_elastic_resample(const float* noise, int64_t D, int64_t H, int64_t W, const float magnitude_rev[3],
                                    const float* in, float* out, int64_t n_volumes, int64_t X, int64_t Y, int64_t Z,
                                    void* stream) {
    ElasticParams P;
    int rc = fill_elastic(P, "skb_elastic_resample", noise, D, H, W, magnitude_rev, X, Y, Z);
    if (rc) return rc;
    SKB_REQUIRE(in && out && in != out && n_volumes >= 1, "skb_elastic_resample: bad argument (in-place is not supported)");
    const long long V = X * Y * Z;
    elastic_resample_kernel<<<(unsigned)((V + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(P, in, out, n_volumes);
    SKB_LAUNCH_CHECK("elastic_resample_kernel");
    return SKB_OK;
}

extern "C" int skb_elastic_points(const float* noise, int64_t D, int64_t H, int64_t W, const float magnitude_rev[3],
                                  const void* points, int points_are_int64, int64_t n_points, int64_t X, int64_t Y,
                                  int64_t Z, void* out_points, void* stream) {
    ElasticParams P;
    int rc = fill_elastic(P, "skb_elastic_points", noise, D, H, W, magnitude_rev, X, Y, Z);
    if (rc) return rc;
    SKB_REQUIRE(n_points >= 0, "skb_elastic_points: bad argument");
    if (n_points == 0) return SKB_OK;
    SKB_REQUIRE(points && out_points, "skb_elastic_points: NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned nb = (unsigned)((n_points + 127) / 128);
    if (points_are_int64)
        elastic_points_kernel<long long><<<nb, 128, 0, st>>>(P, static_cast<const long long*>(points), static_cast<long long*>(out_points), n_points);
    else
        elastic_points_kernel<float><<<nb, 128, 0, st>>>(P, static_cast<const float*>(points), static_cast<float*>(out_points), n_points);
    SKB_LAUNCH_CHECK("elastic_points_kernel");
    return SKB_OK;
}
