// Fused vector -> embedding -> skeleton-label gather (kernel (a) of the north star) and the
// stand-alone forms of its two halves.
//
// Reference arithmetic restated bit-for-bit:
//   skoots/lib/vector_to_embedding.py:79-132   phi = idx + v*s ; N-1 hops {round, clamp to
//                                              [0,dim] (sic), fp32 ravel, clamp, gather, add}
//   skoots/lib/eval.py:258-284                 crop grid, crop-local walk, origin added in fp32
//   skoots/lib/skeleton.py:678-695             round, clamp to [0,dim-1], gather labels, int32
// Every fp32 operation is issued with explicit round-to-nearest intrinsics so nvcc cannot
// contract the reference's separately-rounded multiply and add into an FMA.
//
// Memory plan: each thread owns 8 consecutive voxels of the flat (Z-fastest) index, i.e. one
// 16-byte load per fp16 vector channel and one/two 16-byte stores of labels, all streaming
// (L1::no_allocate).  The label side is read through the bit-packed mask written by the CCL
// tile kernel (1/8 B per voxel, L2-resident over the +-scale window) and, only for targets
// that are foreground, the sparse parent array.  A zero vector (background, ~95 % of a volume)
// short-cuts to "label of myself", which is decided from one byte of the bit mask.
#include <stddef.h>
#include <stdlib.h>

#include "skb_common.cuh"

struct AsmParams {
    const void* vec;        // channel 0 of the field being assembled (batch b)
    const void* vec_hops;   // field the N>1 hops gather from (batch 0 — reference `take` quirk)
    long long cstride;      // elements between channels
    int X, Y, Z;            // global volume
    int Zl, z_off;          // this call covers z in [z_off, z_off+Zl): vec/out are (.,X,Y,Zl) slabs (whole volume: Z, 0)
    const ull* halo_lo;     // slab mode: bits of global word k0-1 / k1 per row (the Z-neighbours' boundary planes)
    const ull* halo_hi;
    int label_halo;         // slab mode: how many planes beyond each face the halo words really describe (0 = not checked)
    unsigned* status;       // slab mode: SKB_STATUS_HALO_RANGE is OR-ed in when a target lies beyond them (may be NULL)
    const unsigned* density;  // probe result: sampled voxels that carry a vector (NULL: no probe, the sparse kernel runs)
    unsigned dense_from;      // the dense kernel instantiation runs when *density >= dense_from, the sparse one otherwise
    int dense_work;           // dense instantiation: work items of a chunk from which every lane resolves its own voxels
    int planar;             // 2-D mode: the volume is a stack (X = slices, Y, Z = image axes) and `vec` is (slices, 2, Y, Z):
    long long plane;        //   two channels per slice, `plane` = Y * Z elements each; the slice axis never moves
    const void* vhalo_lo;   // slab mode, N > 1: where hops that leave the slab inside their crop read the Z-neighbours' vectors:
    const void* vhalo_hi;   //   (3,X,Y,depth) arrays whose LAST vh planes (lo) / FIRST vh planes (hi) are the planes
    int vh;                 //   [z_off - vh, z_off) / [z_off+Zl, z_off+Zl+vh) of the field.  depth = vh: packed copies of the
    int vdepth_lo, vdepth_hi;  // faces; depth = the neighbour's slab depth: the neighbour's own slab, mapped over NVLink
    float s[3];
    int N;
    double decay;
    int cs[3];              // crop size clamped to the volume
    int ov[3];
    int step[3];
    int shifted[3];         // the axis has a shifted-back last crop
    int fast_ok;            // zero-vector voxels provably resolve to themselves
    int vec_aligned;        // 16-byte loads of the vector channels are legal
    int flat_bits;          // Z % 64 == 0: the bit mask has no row padding, bit index == voxel index
    int single_crop;        // the whole volume is one crop with no overlap: owner origin is 0 everywhere
    // label source
    const ull* bits;
    const int* parent;
    int ZW;                 // words per row of `bits` (slab mode: of the slab's compact bit mask, Zl / 64)
    const void* dense;
    int dense_dtype;
};

__device__ __forceinline__ int owner_origin(int g, int dim, int size, int ov, int step, int shifted) {
    // last crop (in the reference's loop order) whose written interior contains g; -1 if none
    if (shifted && g >= dim - size + ov && g < dim - ov) return dim - size;
    if (g >= ov) {
        int o = ((g - ov) / step) * step;
        if (o + size <= dim) return o;
    }
    return -1;
}

template <typename VecT>
__device__ __forceinline__ float load_vec(const void* base, long long idx) {
    return skb_to_float<VecT>(__ldg(static_cast<const VecT*>(base) + idx));
}
template <>
__device__ __forceinline__ float load_vec<__nv_bfloat16>(const void* base, long long idx) {
    unsigned short raw = __ldg(static_cast<const unsigned short*>(base) + idx);
    return __uint_as_float((unsigned)raw << 16);
}

// crop-local walk: (lx,ly,lz) local integer coords, v* own vector, o* crop origin.
// returns the embedding in the crop-local frame (before the origin is added).
template <typename VecT>
__device__ __forceinline__ void walk(const AsmParams& P, int lx, int ly, int lz, int ox, int oy, int oz, float v0,
                                     float v1, float v2, float& mx, float& my, float& mz) {
    mx = __fadd_rn((float)lx, __fmul_rn(v0, P.s[0]));
    my = __fadd_rn((float)ly, __fmul_rn(v1, P.s[1]));
    mz = __fadd_rn((float)lz, __fmul_rn(v2, P.s[2]));
    if (P.N > 1) {
        const float fy = (float)P.cs[1], fz = (float)P.cs[2];
        const long long total = (long long)P.cs[0] * P.cs[1] * P.cs[2];
        const float fmax_index = (float)(total - 1);
        const unsigned ucz = (unsigned)P.cs[2], ucy = (unsigned)P.cs[1];
        double k = 1.0;
        for (int it = 1; it < P.N; ++it) {
            k *= P.decay;
            const float kf = (float)k;
            float ix = fminf(fmaxf(rintf(mx), 0.f), (float)P.cs[0]);
            float iy = fminf(fmaxf(rintf(my), 0.f), fy);
            float iz = fminf(fmaxf(rintf(mz), 0.f), fz);
            float f = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(ix, fy), fz), __fmul_rn(iy, fz)), iz);
            f = fminf(fmaxf(f, 0.f), fmax_index);
            long long fi = (long long)f;
            if (fi > total - 1) fi = total - 1;  // (float)(total-1) may round up; torch would raise here
            unsigned u = (unsigned)fi;
            unsigned qz = u / ucz, cz = u - qz * ucz;
            unsigned cx = qz / ucy, cy = qz - cx * ucy;
            // the field holds z in [z_off, z_off+Zl) (whole volume: [0, Z)); a hop that leaves a slab inside its crop
            // reads the neighbour's planes from the vector halo
            const long long row = (long long)(ox + (int)cx) * P.Y + (oy + (int)cy);
            const int zl = oz + (int)cz - P.z_off;
            const void* hb = P.vec_hops;
            long long g = row * P.Zl + zl, hs = P.cstride;
            if (zl < 0 || zl >= P.Zl) {
                const bool lo = zl < 0;
                const int h = lo ? zl + P.vh : zl - P.Zl;
                hb = lo ? P.vhalo_lo : P.vhalo_hi;
                if (hb == nullptr || h < 0 || h >= P.vh) {  // cannot happen when the halo covers crop - overlap - 1 planes
                    if (P.status) atomicOr(P.status, SKB_STATUS_HALO_RANGE);
                    hb = P.vec_hops; g = row * P.Zl + (lo ? 0 : P.Zl - 1);
                } else {
                    const int depth = lo ? P.vdepth_lo : P.vdepth_hi;
                    g = row * depth + (lo ? depth - P.vh + h : h);
                    hs = (long long)P.X * P.Y * depth;
                }
            }
            float h0 = load_vec<VecT>(hb, g);
            float h1 = load_vec<VecT>(hb, g + hs);
            float h2 = load_vec<VecT>(hb, g + 2 * hs);
            mx = __fadd_rn(mx, __fmul_rn(h0, __fmul_rn(kf, P.s[0])));
            my = __fadd_rn(my, __fmul_rn(h1, __fmul_rn(kf, P.s[1])));
            mz = __fadd_rn(mz, __fmul_rn(h2, __fmul_rn(kf, P.s[2])));
        }
    }
}

__device__ __forceinline__ int clamp_index(float e, int dim) {
    return (int)fminf(fmaxf(rintf(e), 0.f), (float)(dim - 1));
}

__device__ __forceinline__ int label_at(const AsmParams& P, int tx, int ty, int tz) {
    long long rowi = (long long)tx * P.Y + ty;
    if (P.dense) {
        long long t = rowi * P.Z + tz;
        if (P.dense_dtype == SKB_I16) return (int)__ldg(static_cast<const short*>(P.dense) + t);
        if (P.dense_dtype == SKB_I32) return __ldg(static_cast<const int*>(P.dense) + t);
        return (int)__ldg(static_cast<const unsigned char*>(P.dense) + t);
    }
    ull w;
    if (tz < P.z_off) {
        w = P.halo_lo ? __ldg(P.halo_lo + rowi) : 0ull;
        // the neighbour only sent the runs of its last `label_halo` planes: a target beyond them would be answered
        // from a word that does not describe it — report instead of returning a wrong label
        if (P.label_halo && tz < P.z_off - P.label_halo && P.status) atomicOr(P.status, SKB_STATUS_HALO_RANGE);
    } else if (tz >= P.z_off + P.Zl) {
        w = P.halo_hi ? __ldg(P.halo_hi + rowi) : 0ull;
        if (P.label_halo && tz >= P.z_off + P.Zl + P.label_halo && P.status) atomicOr(P.status, SKB_STATUS_HALO_RANGE);
    } else w = __ldg(P.bits + rowi * P.ZW + ((tz - P.z_off) >> 6));
    if (!((w >> (tz & 63)) & 1ull)) return 0;
    return skb_sparse_label(P.parent, (int)(rowi * P.Z + tz));
}

template <typename VecT>
__device__ __forceinline__ int assemble_voxel(const AsmParams& P, int x, int y, int z, float v0, float v1, float v2) {
    int ox = 0, oy = 0, oz = 0;
    if (!P.single_crop) {
        ox = owner_origin(x, P.X, P.cs[0], P.ov[0], P.step[0], P.shifted[0]);
        oy = owner_origin(y, P.Y, P.cs[1], P.ov[1], P.step[1], P.shifted[1]);
        oz = owner_origin(z, P.Z, P.cs[2], P.ov[2], P.step[2], P.shifted[2]);
        if ((ox | oy | oz) < 0) return 0;  // outer margin: never written by the reference (eval.py:259-269)
    }
    float mx, my, mz;
    walk<VecT>(P, x - ox, y - oy, z - oz, ox, oy, oz, v0, v1, v2, mx, my, mz);
    float ex = __fadd_rn(mx, (float)ox), ey = __fadd_rn(my, (float)oy), ez = __fadd_rn(mz, (float)oz);
    return label_at(P, clamp_index(ex, P.X), clamp_index(ey, P.Y), clamp_index(ez, P.Z));
}

// One warp owns 256 consecutive voxels (8 per lane).  Voxels that need real work — a non-zero
// vector, or a zero vector sitting on a foreground voxel — are compacted into a per-warp queue in
// shared memory and processed one per lane, so the divergent part (walk + dependent label reads)
// always runs with full warps and 32 independent load chains in flight.  Everything else is a
// pure stream: 3 x 16-byte loads, one byte of the bit mask, 2 x 16-byte stores, ~40 instructions
// per warp (the zero test works on the raw bit patterns; nothing is converted to float).
constexpr int ASM_WARPS = 8;

template <typename VecT> struct RawOf { typedef unsigned short type; };
template <> struct RawOf<float> { typedef unsigned type; };

template <typename VecT>
__device__ __forceinline__ float raw_to_float(typename RawOf<VecT>::type r);
template <> __device__ __forceinline__ float raw_to_float<__half>(unsigned short r) { return __half2float(__ushort_as_half(r)); }
template <> __device__ __forceinline__ float raw_to_float<__nv_bfloat16>(unsigned short r) { return __uint_as_float((unsigned)r << 16); }
template <> __device__ __forceinline__ float raw_to_float<float>(unsigned r) { return __uint_as_float(r); }

// Raw bits of 8 consecutive elements of one channel, kept as 32-bit words (4 for 16-bit types, 8 for fp32).
template <typename VecT> struct Raw8 {
    static constexpr int NW = sizeof(typename RawOf<VecT>::type) * 2;  // words
    unsigned w[NW];
};

template <typename VecT, bool FULL>
__device__ __forceinline__ Raw8<VecT> load_raw8(const void* base, long long idx, int nvalid) {
    typedef typename RawOf<VecT>::type raw_t;
    Raw8<VecT> r;
    const raw_t* p = static_cast<const raw_t*>(base) + idx;
    if (FULL) {
        uint4 q = skb_ld_stream16(p);
        r.w[0] = q.x; r.w[1] = q.y; r.w[2] = q.z; r.w[3] = q.w;
        if (Raw8<VecT>::NW == 8) {
            uint4 q1 = skb_ld_stream16(reinterpret_cast<const char*>(p) + 16);
            r.w[4] = q1.x; r.w[5] = q1.y; r.w[6] = q1.z; r.w[7] = q1.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < Raw8<VecT>::NW; ++k) r.w[k] = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < nvalid) {
                const unsigned e = (unsigned)__ldg(p + j);
                if (sizeof(raw_t) == 2) r.w[j >> 1] |= e << (16 * (j & 1));
                else r.w[j] = e;
            }
        }
    }
    return r;
}

template <typename VecT>
__device__ __forceinline__ typename RawOf<VecT>::type raw_elem(const unsigned* words, int j) {
    typedef typename RawOf<VecT>::type raw_t;
    if (sizeof(raw_t) == 2) return (raw_t)(words[j >> 1] >> (16 * (j & 1)));
    return (raw_t)words[j];
}

// per-voxel "vector is non-zero" bits (sign ignored: -0 * s adds nothing)
template <typename VecT>
__device__ __forceinline__ unsigned nonzero_bits(const Raw8<VecT>& a, const Raw8<VecT>& b, const Raw8<VecT>& c) {
    unsigned work = 0;
    if (Raw8<VecT>::NW == 4) {
        unsigned o[4], any = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { o[k] = (a.w[k] | b.w[k] | c.w[k]) & 0x7fff7fffu; any |= o[k]; }
        if (any) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                work |= ((unsigned)((o[k] & 0xffffu) != 0u) << (2 * k)) | ((unsigned)((o[k] >> 16) != 0u) << (2 * k + 1));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) work |= (unsigned)(((a.w[j] | b.w[j] | c.w[j]) & 0x7fffffffu) != 0u) << j;
    }
    return work;
}

// DENSE chunks.  The density sweep (bench.py extras) showed the path at 0.24 of its roofline when half of the voxels carry a
// vector against 0.81 on the sparse headline volume: every chunk then takes four rounds through the warp queue (~750
// warp-instructions, three dependent loads per round).  For such chunks it is cheaper to let every lane resolve ITS OWN 8
// voxels, stage by stage with unconditional loads (a voxel without work reads index 0): 8 independent loads in flight
// per lane and stage, no shared-memory traffic, one index division per lane.  A first attempt as a branch inside the one
// kernel cost the SPARSE path its registers (inlined: spills; out of line: a stack frame per chunk — the headline gather
// went from 3.61 to 4.60 ms), so this lives in a second instantiation of the kernel (DENSE = true, 3 CTAs per SM, more
// registers); a probe samples the field's density on the device and both instantiations are launched: the one the probe
// did not pick returns at once (no host decision, no synchronisation).
template <typename VecT>
__device__ __forceinline__ void resolve_own8(const AsmParams& P, const Raw8<VecT>& a, const Raw8<VecT>& b, const Raw8<VecT>& c,
                                             unsigned work, long long i0, unsigned (&lab)[8]) {
    const unsigned q = (unsigned)i0 / (unsigned)P.Zl;
    const int z0 = (int)((unsigned)i0 - q * (unsigned)P.Zl) + P.z_off;  // the 8 voxels share a row (Zl % 8 == 0)
    const int x = (int)(q / (unsigned)P.Y), y = (int)(q - (unsigned)x * (unsigned)P.Y);
    const float fx = (float)x, fy = (float)y;
    const ull* wp[8];
    int tv[8], tb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float ex = __fadd_rn(fx, __fmul_rn(raw_to_float<VecT>(raw_elem<VecT>(a.w, j)), P.s[0]));
        const float ey = __fadd_rn(fy, __fmul_rn(raw_to_float<VecT>(raw_elem<VecT>(b.w, j)), P.s[1]));
        const float ez = __fadd_rn((float)(z0 + j), __fmul_rn(raw_to_float<VecT>(raw_elem<VecT>(c.w, j)), P.s[2]));
        const int tx = clamp_index(ex, P.X), ty = clamp_index(ey, P.Y), tz = clamp_index(ez, P.Z);
        const long long rowi = (long long)tx * P.Y + ty;
        const bool have = ((work >> j) & 1u) != 0u;
        wp[j] = have ? P.bits + rowi * P.ZW + (tz >> 6) : P.bits;
        tv[j] = have ? (int)(rowi * P.Z + tz) : -1;
        tb[j] = tz & 63;
    }
    ull ww[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ww[j] = __ldg(wp[j]);
    int p1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (!((ww[j] >> tb[j]) & 1ull)) tv[j] = -1;
        p1[j] = __ldg(P.parent + (tv[j] < 0 ? 0 : tv[j]));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int p2 = __ldg(P.parent + ((tv[j] < 0 || p1[j] < 0) ? 0 : p1[j]));  // parent[0] itself may be uninitialised
        const int code = p1[j] < 0 ? p1[j] : p2;
        lab[j] = tv[j] < 0 ? 0u : (unsigned)(-code);
    }
}

// work items of a 256-voxel chunk from which the per-lane form is taken.  Measured on 1024x1024x256 (whole pass, ms) for
// thresholds 16 / 24 / 40 / 64 / never: 25 % dense 1.36 / 1.35 / 1.33 / 1.32 / 1.50, 51 % dense 1.40 / 1.39 / 1.39 / 1.39 / 2.07
constexpr int ASM_DENSE_WORK = 48;

// What one lane holds for its 8 voxels between the load and the processing of a 256-voxel chunk.
template <typename VecT> struct ChunkRegs {
    Raw8<VecT> r0, r1, r2;
    unsigned self;  // foreground bits of my own 8 voxels (zero-vector voxels resolve to themselves)
};

// FULL = all 32 lanes own 8 valid voxels and 16-byte loads are legal: no per-element guards at all.
template <typename VecT, bool FULL>
__device__ __forceinline__ ChunkRegs<VecT> load_chunk(const AsmParams& P, long long V, long long warp_base) {
    const int lane = threadIdx.x & 31;
    const long long i0 = warp_base + lane * 8;
    const long long left = V - i0;
    const int nvalid = FULL ? 8 : (left >= 8 ? 8 : (left > 0 ? (int)left : 0));
    const unsigned uz = (unsigned)P.Zl;
    ChunkRegs<VecT> c;
    // every independent load is issued before anything waits on one of them
    if (P.planar) {
        // (slices, 2, Y, Z): the slice's two channels are the y / z components; the x (slice) component is zero
#pragma unroll
        for (int k = 0; k < Raw8<VecT>::NW; ++k) c.r0.w[k] = 0u;
        if (FULL) {  // plane % 8 == 0: the 8 voxels share a slice
            const long long sl = i0 / P.plane, a0 = sl * 2 * P.plane + (i0 - sl * P.plane);
            c.r1 = load_raw8<VecT, true>(P.vec, a0, 8);
            c.r2 = load_raw8<VecT, true>(P.vec, a0 + P.plane, 8);
        } else {
            typedef typename RawOf<VecT>::type raw_t;
#pragma unroll
            for (int k = 0; k < Raw8<VecT>::NW; ++k) { c.r1.w[k] = 0u; c.r2.w[k] = 0u; }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < nvalid) {
                    const long long i = i0 + j, sl = i / P.plane, a = sl * 2 * P.plane + (i - sl * P.plane);
                    const unsigned e1 = (unsigned)__ldg(static_cast<const raw_t*>(P.vec) + a);
                    const unsigned e2 = (unsigned)__ldg(static_cast<const raw_t*>(P.vec) + a + P.plane);
                    if (sizeof(raw_t) == 2) { c.r1.w[j >> 1] |= e1 << (16 * (j & 1)); c.r2.w[j >> 1] |= e2 << (16 * (j & 1)); }
                    else { c.r1.w[j] = e1; c.r2.w[j] = e2; }
                }
            }
        }
    } else {
        c.r0 = load_raw8<VecT, FULL>(P.vec, i0, nvalid);
        c.r1 = load_raw8<VecT, FULL>(P.vec, i0 + P.cstride, nvalid);
        c.r2 = load_raw8<VecT, FULL>(P.vec, i0 + 2 * P.cstride, nvalid);
    }
    c.self = 0;
    if (P.fast_ok && !P.dense && nvalid > 0) {
        if (P.flat_bits) {  // no row padding: bit index == (slab-local) voxel index
            c.self = (unsigned)__ldg(reinterpret_cast<const unsigned char*>(P.bits) + ((size_t)i0 >> 3));
        } else {
            unsigned q = (unsigned)i0 / uz;
            int z = (int)((unsigned)i0 - q * uz);
            if (z + 8 <= P.Z) {
                const long long wi = (long long)q * P.ZW + (z >> 6);
                const int sh = z & 63;
                ull w = __ldg(P.bits + wi) >> sh;
                if (sh > 56) w |= __ldg(P.bits + wi + 1) << (64 - sh);
                c.self = (unsigned)(w & 0xFFull);
            } else {
                for (int j = 0; j < nvalid; ++j) {
                    ull w = __ldg(P.bits + (long long)q * P.ZW + (z >> 6));
                    c.self |= (unsigned)((w >> (z & 63)) & 1ull) << j;
                    if (++z == P.Z) { z = 0; ++q; }
                }
            }
        }
    }
    return c;
}

template <typename VecT, typename OutT, bool FULL, bool DENSE = false>
__device__ __forceinline__ void process_chunk(const AsmParams& P, const ChunkRegs<VecT>& c, OutT* __restrict__ out,
                                              long long V, long long warp_base,
                                              unsigned (*s_raw)[32][Raw8<VecT>::NW], int* s_res, unsigned char* s_queue) {
    const int lane = threadIdx.x & 31;
    const long long i0 = warp_base + lane * 8;
    const long long left = V - i0;
    const int nvalid = FULL ? 8 : (left >= 8 ? 8 : (left > 0 ? (int)left : 0));
    const unsigned uz = (unsigned)P.Zl, uy = (unsigned)P.Y;  // index decomposition runs over the slab
    const unsigned valid_mask = FULL ? 0xFFu : ((1u << nvalid) - 1u);
    unsigned work;
    if (P.fast_ok && !P.dense) work = (nonzero_bits<VecT>(c.r0, c.r1, c.r2) | c.self) & valid_mask;
    else work = valid_mask;  // N>1 over a >2^24-voxel crop, or a dense label volume (compatibility path)

    unsigned lab[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) lab[j] = 0u;

    if (__ballot_sync(0xffffffffu, work != 0u) != 0u) {
        // warp-wide compaction of the work items
        const int cnt = __popc(work);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (DENSE && total >= P.dense_work) {  // warp-uniform; the dense instantiation runs whole-volume N = 1 passes only
            if (work) resolve_own8<VecT>(P, c.r0, c.r1, c.r2, work, i0, lab);
        } else {
        if (work) {
            int at = incl - cnt;
            for (unsigned m = work; m; m &= m - 1) s_queue[at++] = (unsigned char)((__ffs((int)m) - 1) * 32 + lane);
#pragma unroll
            for (int k = 0; k < Raw8<VecT>::NW; ++k) {
                s_raw[0][lane][k] = c.r0.w[k];
                s_raw[1][lane][k] = c.r1.w[k];
                s_raw[2][lane][k] = c.r2.w[k];
            }
        }
        __syncwarp();
        for (int qi = lane; qi < total; qi += 32) {
            const int code = s_queue[qi];
            const int src = code & 31, j = code >> 5;
            const unsigned vi = (unsigned)(warp_base + src * 8 + j);
            const unsigned q = vi / uz;
            const int z = (int)(vi - q * uz) + P.z_off;
            const int x = (int)(q / uy);
            const int y = (int)(q - (unsigned)x * uy);
            s_res[code] = assemble_voxel<VecT>(P, x, y, z, raw_to_float<VecT>(raw_elem<VecT>(s_raw[0][src], j)),
                                               raw_to_float<VecT>(raw_elem<VecT>(s_raw[1][src], j)),
                                               raw_to_float<VecT>(raw_elem<VecT>(s_raw[2][src], j)));
        }
        __syncwarp();
        if (work) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int v = s_res[j * 32 + lane];  // unconditional read, selected below: keeps lab[] in registers
                lab[j] = ((work >> j) & 1u) ? (unsigned)v : 0u;
            }
        }
        __syncwarp();  // the warp's shared-memory scratch is reused by its next chunk
        }
    }

    if (FULL || nvalid == 8) {
        if (sizeof(OutT) == 2) {
            skb_st_stream16(out + i0, make_uint4((lab[0] & 0xffffu) | (lab[1] << 16), (lab[2] & 0xffffu) | (lab[3] << 16),
                                                 (lab[4] & 0xffffu) | (lab[5] << 16), (lab[6] & 0xffffu) | (lab[7] << 16)));
        } else {
            skb_st_stream16(out + i0, make_uint4(lab[0], lab[1], lab[2], lab[3]));
            skb_st_stream16(out + i0 + 4, make_uint4(lab[4], lab[5], lab[6], lab[7]));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < nvalid) out[i0 + j] = (OutT)lab[j];
    }
}

// Main kernel: PERSISTENT warps over the full, aligned 256-voxel chunks [0, n_chunks).  A warp issues the
// streaming loads of its next chunk before it starts on the current one, so a warp that is busy with
// the latency-bound part (compaction, dependent label reads) still keeps 1.5 KB of loads in flight.
template <typename VecT, typename OutT, bool DENSE>
__global__ void __launch_bounds__(32 * ASM_WARPS, DENSE ? 3 : 4) assemble_kernel(AsmParams P, OutT* __restrict__ out, long long V,
                                                                                  unsigned chunk_begin, unsigned n_chunks) {
    __shared__ unsigned s_raw[ASM_WARPS][3][32][Raw8<VecT>::NW];
    __shared__ int s_res[ASM_WARPS][256];
    __shared__ unsigned char s_queue[ASM_WARPS][256];
    if (P.density && ((*P.density >= P.dense_from) != DENSE)) return;  // the probe picked the other instantiation
    const int warp = threadIdx.x >> 5;
    const unsigned stride = gridDim.x * ASM_WARPS;
    unsigned c = chunk_begin + blockIdx.x * ASM_WARPS + warp;  // chunks [chunk_begin, n_chunks)
    if (c >= n_chunks) return;
    ChunkRegs<VecT> cur = load_chunk<VecT, true>(P, V, (long long)c * 256);
    for (;;) {
        const unsigned cn = c + stride;
        ChunkRegs<VecT> nxt = cur;
        if (cn < n_chunks) nxt = load_chunk<VecT, true>(P, V, (long long)cn * 256);
        process_chunk<VecT, OutT, true, DENSE>(P, cur, out, V, (long long)c * 256, s_raw[warp], s_res[warp], s_queue[warp]);
        if (cn >= n_chunks) break;
        c = cn;
        cur = nxt;
    }
}

// ------------------------------------------------------------------------------------------
// TMA-staged form of the main kernel (the default whenever the field is aligned, the bit mask has no row
// padding and zero vectors resolve to themselves — the headline mode).
//
// ncu on the register-prefetch kernel above: 57 % long-scoreboard stalls, 30 % of all samples on the first
// use of the prefetched chunk — one chunk (1.5 KB) in flight per warp is not enough while the warp is
// held up by the label look-ups of its current chunk.  Here every warp owns a RING of stages in shared
// memory; lane 0 fills a stage with three bulk asynchronous copies (cp.async.bulk global -> shared, one
// per vector channel, completion counted in bytes on the stage's mbarrier), so the loads of the next
// STAGES-1 chunks are in flight without holding a single register, whatever the warp is doing.  The
// queue processing reads other lanes' vectors straight from the stage (no second copy in shared memory).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(ull* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(ull* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, ull* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(ull* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

template <typename VecT> struct StageCfg {
    typedef typename RawOf<VecT>::type raw_t;
    static constexpr int CH_BYTES = 256 * (int)sizeof(raw_t);           // one channel of one 256-voxel chunk
    static constexpr int STAGES = sizeof(raw_t) == 2 ? 3 : 2;
    static constexpr int STAGE_BYTES = 3 * CH_BYTES + 32;  // three channels + the chunk's 256 foreground bits
    static constexpr int WARP_BYTES = STAGES * STAGE_BYTES + 256 * 4 + 256 + 64;  // ring + results + queue + barriers
};

__device__ __forceinline__ void ldgsts16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ldgsts_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void ldgsts_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// BULK = true : lane 0 fills a stage with three cp.async.bulk copies (TMA engine, mbarrier completion);
// BULK = false: every lane copies its own 16 bytes per channel with cp.async (LDGSTS), one commit group per stage.
template <typename VecT, typename OutT, bool BULK>
__global__ void __launch_bounds__(32 * ASM_WARPS, 4) assemble_tma_kernel(AsmParams P, OutT* __restrict__ out,
                                                                          unsigned chunk_begin, unsigned n_chunks) {
    typedef StageCfg<VecT> C;
    typedef typename C::raw_t raw_t;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* base = smem + (size_t)warp * C::WARP_BYTES;
    unsigned char* ring = base;
    int* s_res = reinterpret_cast<int*>(base + C::STAGES * C::STAGE_BYTES);
    unsigned char* s_queue = reinterpret_cast<unsigned char*>(s_res + 256);
    ull* bars = reinterpret_cast<ull*>(s_queue + 256);
    const unsigned stride = gridDim.x * ASM_WARPS;
    const unsigned first = chunk_begin + blockIdx.x * ASM_WARPS + warp;
    if (first >= n_chunks) return;
    const unsigned uz = (unsigned)P.Zl, uy = (unsigned)P.Y;

    auto fill = [&](int stage, unsigned c) {  // BULK: lane 0 only; else every lane
        const raw_t* src = static_cast<const raw_t*>(P.vec) + (size_t)c * 256;
        const unsigned char* bsrc = reinterpret_cast<const unsigned char*>(P.bits) + (size_t)c * 32;  // no row padding: bit = voxel
        unsigned char* dst = ring + (size_t)stage * C::STAGE_BYTES;
        if (BULK) {
            ull* bar = bars + stage;
            mbar_expect_tx(bar, (unsigned)C::STAGE_BYTES);
            bulk_load(dst, src, C::CH_BYTES, bar);
            bulk_load(dst + C::CH_BYTES, src + P.cstride, C::CH_BYTES, bar);
            bulk_load(dst + 2 * C::CH_BYTES, src + 2 * P.cstride, C::CH_BYTES, bar);
            bulk_load(dst + 3 * C::CH_BYTES, bsrc, 32u, bar);
        } else {
            if (lane < 2) ldgsts16(dst + 3 * C::CH_BYTES + lane * 16, bsrc + lane * 16);
            constexpr int PER_LANE = 8 * (int)sizeof(raw_t);  // bytes of a lane's 8 voxels in one channel: 16 or 32
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const unsigned char* g = reinterpret_cast<const unsigned char*>(src + (size_t)ch * P.cstride) + lane * PER_LANE;
                unsigned char* d = dst + ch * C::CH_BYTES + lane * PER_LANE;
#pragma unroll
                for (int b = 0; b < PER_LANE; b += 16) ldgsts16(d + b, g + b);
            }
        }
    };

    if (BULK) {
        if (lane == 0) {
            for (int st = 0; st < C::STAGES; ++st) mbar_init(bars + st, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int st = 0; st < C::STAGES; ++st) {
                const unsigned long long c = (unsigned long long)first + (unsigned long long)st * stride;
                if (c < n_chunks) fill(st, (unsigned)c);
            }
        }
    } else {
        for (int st = 0; st < C::STAGES; ++st) {
            const unsigned long long c = (unsigned long long)first + (unsigned long long)st * stride;
            if (c < n_chunks) fill(st, (unsigned)c);
            ldgsts_commit();  // one group per stage, empty or not: the group count stays uniform
        }
    }
    __syncwarp();

    unsigned it = 0;
    for (unsigned c = first;; ++it) {
        const int stage = (int)(it % C::STAGES);
        const unsigned parity = (it / C::STAGES) & 1u;
        const long long warp_base = (long long)c * 256;
        const long long i0 = warp_base + lane * 8;
        if (BULK) {
            mbar_wait(bars + stage, parity);
        } else {
            ldgsts_wait<C::STAGES - 1>();  // my own copies of the oldest stage have landed ...
            __syncwarp();                  // ... and so have every other lane's
        }
        const raw_t* ch0 = reinterpret_cast<const raw_t*>(ring + (size_t)stage * C::STAGE_BYTES);
        const raw_t* ch1 = ch0 + 256;
        const raw_t* ch2 = ch1 + 256;
        const unsigned self = reinterpret_cast<const unsigned char*>(ch2 + 256)[lane];  // foreground bits of my own 8 voxels
        Raw8<VecT> r0, r1, r2;
        {
            const uint4* q0 = reinterpret_cast<const uint4*>(ch0 + lane * 8);
            const uint4* q1 = reinterpret_cast<const uint4*>(ch1 + lane * 8);
            const uint4* q2 = reinterpret_cast<const uint4*>(ch2 + lane * 8);
#pragma unroll
            for (int h = 0; h < Raw8<VecT>::NW / 4; ++h) {
                const uint4 a = q0[h], b = q1[h], d = q2[h];
                r0.w[4 * h] = a.x; r0.w[4 * h + 1] = a.y; r0.w[4 * h + 2] = a.z; r0.w[4 * h + 3] = a.w;
                r1.w[4 * h] = b.x; r1.w[4 * h + 1] = b.y; r1.w[4 * h + 2] = b.z; r1.w[4 * h + 3] = b.w;
                r2.w[4 * h] = d.x; r2.w[4 * h + 1] = d.y; r2.w[4 * h + 2] = d.z; r2.w[4 * h + 3] = d.w;
            }
        }
        const unsigned work = nonzero_bits<VecT>(r0, r1, r2) | self;
        unsigned lab[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) lab[j] = 0u;
        if (__ballot_sync(0xffffffffu, work != 0u) != 0u) {
            const int cnt = __popc(work);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            int at = incl - cnt;
            for (unsigned m = work; m; m &= m - 1) s_queue[at++] = (unsigned char)((__ffs((int)m) - 1) * 32 + lane);
            __syncwarp();
            for (int qi = lane; qi < total; qi += 32) {
                const int code = s_queue[qi];
                const int src = code & 31, j = code >> 5;
                const unsigned vi = (unsigned)(warp_base + src * 8 + j);
                const unsigned q = vi / uz;
                const int z = (int)(vi - q * uz) + P.z_off;
                const int x = (int)(q / uy);
                const int y = (int)(q - (unsigned)x * uy);
                s_res[code] = assemble_voxel<VecT>(P, x, y, z, raw_to_float<VecT>(ch0[src * 8 + j]),
                                                   raw_to_float<VecT>(ch1[src * 8 + j]), raw_to_float<VecT>(ch2[src * 8 + j]));
            }
            __syncwarp();
            if (work) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int v = s_res[j * 32 + lane];
                    lab[j] = ((work >> j) & 1u) ? (unsigned)v : 0u;
                }
            }
        }
        __syncwarp();  // every lane is done with this stage (and with the scratch) before it is refilled
        const unsigned long long cn = (unsigned long long)c + (unsigned long long)C::STAGES * stride;
        if (BULK) {
            if (lane == 0 && cn < n_chunks) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads before the async-proxy overwrite
                fill(stage, (unsigned)cn);
            }
        } else {
            if (cn < n_chunks) fill(stage, (unsigned)cn);
            ldgsts_commit();
        }
        if (sizeof(OutT) == 2) {
            skb_st_stream16(out + i0, make_uint4((lab[0] & 0xffffu) | (lab[1] << 16), (lab[2] & 0xffffu) | (lab[3] << 16),
                                                 (lab[4] & 0xffffu) | (lab[5] << 16), (lab[6] & 0xffffu) | (lab[7] << 16)));
        } else {
            skb_st_stream16(out + i0, make_uint4(lab[0], lab[1], lab[2], lab[3]));
            skb_st_stream16(out + i0 + 4, make_uint4(lab[4], lab[5], lab[6], lab[7]));
        }
        const unsigned long long nx = (unsigned long long)c + stride;
        if (nx >= n_chunks) break;
        c = (unsigned)nx;
    }
}

// Density probe: every thread looks at one 8-voxel group (three 16-byte loads), groups spread evenly over the range; the
// number of sampled voxels that carry a vector is added to *count (zeroed by the host-side memset before the launch).
constexpr unsigned ASM_PROBE_GROUPS = 16384;  // 131 072 sampled voxels
template <typename VecT>
__global__ void __launch_bounds__(256) assemble_probe_kernel(AsmParams P, unsigned chunk_begin, unsigned n_chunks, unsigned* count) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned n = 0;
    if (g < ASM_PROBE_GROUPS) {
        const unsigned long long groups = (unsigned long long)(n_chunks - chunk_begin) * 32ull;
        const long long i0 = ((long long)chunk_begin * 32 + (long long)(groups * g / ASM_PROBE_GROUPS)) * 8;
        const Raw8<VecT> a = load_raw8<VecT, true>(P.vec, i0, 8), b = load_raw8<VecT, true>(P.vec, i0 + P.cstride, 8),
                         c = load_raw8<VecT, true>(P.vec, i0 + 2 * P.cstride, 8);
        n = (unsigned)__popc(nonzero_bits<VecT>(a, b, c));
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, n);
}

// Everything the main kernel does not cover: the ragged last chunk, or all chunks of an unaligned field.
template <typename VecT, typename OutT>
__global__ void __launch_bounds__(32 * ASM_WARPS) assemble_tail_kernel(AsmParams P, OutT* __restrict__ out, long long V,
                                                                       long long first_voxel) {
    __shared__ unsigned s_raw[ASM_WARPS][3][32][Raw8<VecT>::NW];
    __shared__ int s_res[ASM_WARPS][256];
    __shared__ unsigned char s_queue[ASM_WARPS][256];
    const int warp = threadIdx.x >> 5;
    const long long warp_base = first_voxel + ((long long)blockIdx.x * ASM_WARPS + warp) * 256;
    if (warp_base >= V) return;
    const ChunkRegs<VecT> c = load_chunk<VecT, false>(P, V, warp_base);
    process_chunk<VecT, OutT, false>(P, c, out, V, warp_base, s_raw[warp], s_res[warp], s_queue[warp]);
}

// ------------------------------------------------------------------------------------------
// Split form of the N = 1, whole-volume-is-one-crop gather ("stream / resolve").
//
// The fused kernel above cannot start before the labelling has finished.  But 96 % of the voxels
// never look at a label: their vector is zero and they are not skeleton themselves.  So the gather is
// split in two:
//   assemble_stream_kernel   needs only the bit mask.  Pure stream over the vector field (6 B/voxel in);
//                            an 8-voxel group (one lane's 16-byte load per channel) with no work gets
//                            its zeros stored (4 B/voxel out); a group with work is left alone and
//                            recorded as one bit of the chunk's 32-bit flag word.  It runs on a second
//                            stream WHILE the latency-bound labelling chain (tile union-find, boundary
//                            unions, exchanges, merge) runs — HBM-bound next to latency-bound.
//   assemble_resolve_kernel  after the labels exist: walks the flag words, re-reads the vectors of the
//                            flagged groups only (~4 % of the volume), resolves their labels and stores them.
// Non-persistent CTAs on purpose: slots free up continuously, so the (higher-priority) labelling
// kernels always find SM resources next to the stream.
// ------------------------------------------------------------------------------------------
constexpr int STREAM_CHUNKS_PER_WARP = 2;

// The grid is either one CTA per 16 chunks, or — `persistent` — a fixed, small number of CTAs per SM whose warps
// stride over the chunks: the stream then occupies a bounded slice of every SM (threads, registers) for its whole
// duration and the labelling kernels on the other stream always find room next to it.
template <typename VecT, typename OutT>
__global__ void __launch_bounds__(32 * ASM_WARPS) assemble_stream_kernel(AsmParams P, OutT* __restrict__ out,
                                                                         unsigned* __restrict__ flags, unsigned n_chunks) {
    const int lane = threadIdx.x & 31;
    const unsigned stride = gridDim.x * ASM_WARPS * STREAM_CHUNKS_PER_WARP;
    for (unsigned c0 = (blockIdx.x * ASM_WARPS + (threadIdx.x >> 5)) * STREAM_CHUNKS_PER_WARP; c0 < n_chunks; c0 += stride) {
        ChunkRegs<VecT> r[STREAM_CHUNKS_PER_WARP];
#pragma unroll
        for (int u = 0; u < STREAM_CHUNKS_PER_WARP; ++u)
            if (c0 + u < n_chunks) r[u] = load_chunk<VecT, true>(P, 0, (long long)(c0 + u) * 256);
#pragma unroll
        for (int u = 0; u < STREAM_CHUNKS_PER_WARP; ++u) {
            if (c0 + u >= n_chunks) break;  // warp-uniform
            const bool work = (nonzero_bits<VecT>(r[u].r0, r[u].r1, r[u].r2) | r[u].self) != 0u;
            const unsigned m = __ballot_sync(0xffffffffu, work);
            if (!work) {
                OutT* o = out + (long long)(c0 + u) * 256 + lane * 8;
                skb_st_stream16(o, make_uint4(0u, 0u, 0u, 0u));
                if (sizeof(OutT) == 4) skb_st_stream16(o + 4, make_uint4(0u, 0u, 0u, 0u));
            }
            if (lane == 0) flags[c0 + u] = m;
        }
    }
}

// one lane = one flagged 8-voxel group.  The (up to) 8 label look-ups are three dependent loads each (bit-mask
// word, parent, parent of the tile root); they are issued stage by stage for all 8 voxels with UNCONDITIONAL
// loads (a voxel without work reads index 0), so a lane has 8 loads in flight per stage instead of walking the
// voxels one after the other through branches.
template <typename VecT, typename OutT>
__device__ __forceinline__ void resolve_group(const AsmParams& P, long long i0, OutT* __restrict__ out) {
    const Raw8<VecT> a = load_raw8<VecT, true>(P.vec, i0, 8);
    const Raw8<VecT> b = load_raw8<VecT, true>(P.vec, i0 + P.cstride, 8);
    const Raw8<VecT> c = load_raw8<VecT, true>(P.vec, i0 + 2 * P.cstride, 8);
    const unsigned self = (unsigned)__ldg(reinterpret_cast<const unsigned char*>(P.bits) + ((size_t)i0 >> 3));
    const unsigned work = nonzero_bits<VecT>(a, b, c) | self;
    const unsigned q = (unsigned)i0 / (unsigned)P.Zl;
    const int z0 = (int)((unsigned)i0 - q * (unsigned)P.Zl) + P.z_off;  // the 8 voxels share a row (Zl % 8 == 0)
    const int x = (int)(q / (unsigned)P.Y), y = (int)(q - (unsigned)x * (unsigned)P.Y);
    const float fx = (float)x, fy = (float)y;
    const int z_end = P.z_off + P.Zl;
    // stage 1: targets -> address of the 64-bit word that holds the target's bit (slab, or a neighbour's halo plane)
    const ull* wp[8];
    int tv[8], tb[8];
    bool beyond = false;  // a target past the planes the neighbours' runs describe (slab mode): reported, not mislabelled
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float ex = __fadd_rn(fx, __fmul_rn(raw_to_float<VecT>(raw_elem<VecT>(a.w, j)), P.s[0]));
        const float ey = __fadd_rn(fy, __fmul_rn(raw_to_float<VecT>(raw_elem<VecT>(b.w, j)), P.s[1]));
        const float ez = __fadd_rn((float)(z0 + j), __fmul_rn(raw_to_float<VecT>(raw_elem<VecT>(c.w, j)), P.s[2]));
        const int tx = clamp_index(ex, P.X), ty = clamp_index(ey, P.Y), tz = clamp_index(ez, P.Z);
        const long long rowi = (long long)tx * P.Y + ty;
        const ull* p = P.bits + rowi * P.ZW + ((tz - P.z_off) >> 6);
        bool have = ((work >> j) & 1u) != 0u;
        if (tz < P.z_off) {
            beyond |= have && P.label_halo && tz < P.z_off - P.label_halo;
            p = P.halo_lo + rowi; have = have && P.halo_lo != nullptr;
        } else if (tz >= z_end) {
            beyond |= have && P.label_halo && tz >= z_end + P.label_halo;
            p = P.halo_hi + rowi; have = have && P.halo_hi != nullptr;
        }
        wp[j] = have ? p : P.bits;
        tv[j] = have ? (int)(rowi * P.Z + tz) : -1;
        tb[j] = tz & 63;
    }
    if (beyond && P.status) atomicOr(P.status, SKB_STATUS_HALO_RANGE);
    // stage 2: bit-mask words
    ull ww[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ww[j] = __ldg(wp[j]);
    // stage 3: parent of the foreground targets
    int p1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (!((ww[j] >> tb[j]) & 1ull)) tv[j] = -1;
        p1[j] = __ldg(P.parent + (tv[j] < 0 ? 0 : tv[j]));
    }
    // stage 4: a target that is not a tile root holds the tile root's index; the root holds the (negative) label code
    unsigned lab[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int p2 = __ldg(P.parent + ((tv[j] < 0 || p1[j] < 0) ? 0 : p1[j]));  // parent[0] itself may be uninitialised
        const int code = p1[j] < 0 ? p1[j] : p2;
        lab[j] = tv[j] < 0 ? 0u : (unsigned)(-code);
    }
    if (sizeof(OutT) == 2) {
        skb_st_stream16(out + i0, make_uint4((lab[0] & 0xffffu) | (lab[1] << 16), (lab[2] & 0xffffu) | (lab[3] << 16),
                                             (lab[4] & 0xffffu) | (lab[5] << 16), (lab[6] & 0xffffu) | (lab[7] << 16)));
    } else {
        skb_st_stream16(out + i0, make_uint4(lab[0], lab[1], lab[2], lab[3]));
        skb_st_stream16(out + i0 + 4, make_uint4(lab[4], lab[5], lab[6], lab[7]));
    }
}

// A warp takes 32 flag words (32 chunks = 8192 voxels) at a time, lists their flagged groups in shared
// memory and works through the list 32 groups at a time, so the look-ups run with full warps however
// thinly the groups are spread over the chunks.
template <typename VecT, typename OutT>
__global__ void __launch_bounds__(32 * ASM_WARPS) assemble_resolve_kernel(AsmParams P, OutT* __restrict__ out,
                                                                          const unsigned* __restrict__ flags,
                                                                          unsigned n_chunks) {
    __shared__ unsigned short s_list[ASM_WARPS][1024];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned short* list = s_list[warp];
    const unsigned n_blocks = (n_chunks + 31u) / 32u, n_warps = gridDim.x * ASM_WARPS;
    for (unsigned blk = blockIdx.x * ASM_WARPS + warp; blk < n_blocks; blk += n_warps) {
        const unsigned c = blk * 32u + lane;
        const unsigned m = c < n_chunks ? __ldg(flags + c) : 0u;
        const int cnt = __popc(m);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        int at = incl - cnt;
        for (unsigned mm = m; mm; mm &= mm - 1) list[at++] = (unsigned short)((lane << 5) | (__ffs((int)mm) - 1));
        __syncwarp();
        for (int base = 0; base < total; base += 32) {
            const int idx = base + lane;
            if (idx < total) {
                const unsigned e = list[idx];
                const long long group = ((long long)blk * 32 + (e >> 5)) * 32 + (e & 31u);
                resolve_group<VecT, OutT>(P, group * 8, out);
            }
        }
        __syncwarp();  // the list is reused by this warp's next block
    }
}

// ---- stand-alone a1: materialise the embedding ---------------------------------------------------
template <typename VecT>
__global__ void __launch_bounds__(256) vec_embed3d_kernel(AsmParams P, float* __restrict__ out, long long V) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const unsigned uz = (unsigned)P.Z, uy = (unsigned)P.Y;
    unsigned q = (unsigned)i / uz;
    int z = (int)((unsigned)i - q * uz);
    int x = (int)(q / uy);
    int y = (int)(q - (unsigned)x * uy);
    float v0 = load_vec<VecT>(P.vec, i), v1 = load_vec<VecT>(P.vec, i + P.cstride),
          v2 = load_vec<VecT>(P.vec, i + 2 * P.cstride);
    float mx, my, mz;
    walk<VecT>(P, x, y, z, 0, 0, 0, v0, v1, v2, mx, my, mz);
    out[i] = mx;
    out[i + V] = my;
    out[i + 2 * V] = mz;
}

template <typename VecT>
__global__ void __launch_bounds__(256) vec_embed2d_kernel(const VecT* __restrict__ vec, int X, int Y, float s0, float s1,
                                                         float* __restrict__ out, long long total) {
    // total = B*X*Y ; vec/out are (B,2,X,Y)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long plane = (long long)X * Y;
    const long long b = i / plane, r = i - b * plane;
    const int x = (int)(r / Y), y = (int)(r - (long long)x * Y);
    const long long at = b * 2 * plane + r;
    out[at] = __fadd_rn((float)x, __fmul_rn(skb_to_float<VecT>(vec[at]), s0));
    out[at + plane] = __fadd_rn((float)y, __fmul_rn(skb_to_float<VecT>(vec[at + plane]), s1));
}

// N = 1 forms, 8 consecutive elements of the flat index per thread: one 16-byte load (two for fp32) per channel and
// two 16-byte stores per channel, instead of a 2-byte load and a 4-byte store per thread.  `inner` = length of the
// fastest axis, `mid` = length of the axis before it; C = 3 (x, y, z) or 2 (x, y).  The channel plane must be a
// multiple of 8 elements (16-byte aligned planes); a group may run over a row end (Z = 20 in the training crops), so the
// coordinates of the first element are decomposed once and carried for the other seven.
template <typename VecT, int C>
__global__ void __launch_bounds__(256) vec_embed_n1_vec8_kernel(const void* __restrict__ vec, float* __restrict__ out,
                                                                long long per_channel, unsigned inner, unsigned mid,
                                                                float s0, float s1, float s2, long long groups_per_batch,
                                                                long long total_groups) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_groups) return;
    const long long b = g / groups_per_batch;
    const long long i0 = (g - b * groups_per_batch) * 8;  // element index inside one channel of batch b
    unsigned q = (unsigned)(i0 / inner);
    unsigned last = (unsigned)(i0 - (long long)q * inner);
    unsigned c0 = C == 3 ? q / mid : q, c1 = C == 3 ? q - (q / mid) * mid : 0u;
    float idx[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (C == 3) { idx[0][j] = (float)c0; idx[1][j] = (float)c1; idx[2][j] = (float)last; }
        else { idx[0][j] = (float)c0; idx[1][j] = (float)last; idx[2][j] = 0.f; }
        if (++last == inner) {
            last = 0;
            if (C == 3) { if (++c1 == mid) { c1 = 0; ++c0; } }
            else ++c0;
        }
    }
    const float sc[3] = {s0, s1, s2};
    const long long base = b * C * per_channel + i0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const Raw8<VecT> r = load_raw8<VecT, true>(vec, base + (long long)c * per_channel, 8);
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            o[j] = __fadd_rn(idx[c][j], __fmul_rn(raw_to_float<VecT>(raw_elem<VecT>(r.w, j)), sc[c]));
        float* dst = out + base + (long long)c * per_channel;
        skb_st_stream16(dst, make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3])));
        skb_st_stream16(dst + 4, make_uint4(__float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7])));
    }
}

template <typename VecT, int C>
static void launch_vec_embed_n1_vec8(const void* vec, float* out, long long B, long long per_channel, unsigned inner, unsigned mid,
                                     const float* scale, cudaStream_t st) {
    const long long gpb = per_channel / 8, total = B * gpb;
    vec_embed_n1_vec8_kernel<VecT, C><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(vec, out, per_channel, inner, mid, scale[0], scale[1],
                                                                                       C == 3 ? scale[2] : 0.f, gpb, total);
}

template <typename VecT>
__global__ void __launch_bounds__(256) vec_embed_bwd_kernel(const float* __restrict__ go, VecT* __restrict__ gv,
                                                           long long inner, int C, float s0, float s1, float s2,
                                                           long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)((i / inner) % C);
    const float s = c == 0 ? s0 : (c == 1 ? s1 : s2);
    gv[i] = skb_from_float<VecT>(__fmul_rn(go[i], s));
}

// ---- stand-alone a2: gather labels by a materialised embedding -----------------------------------------
template <typename LabT>
__global__ void __launch_bounds__(256) index_by_embed_kernel(const LabT* __restrict__ labels, int Xs, int Ys, int Zs,
                                                            const float* __restrict__ embed, long long n,
                                                            int* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int tx = clamp_index(embed[i], Xs), ty = clamp_index(embed[i + n], Ys), tz = clamp_index(embed[i + 2 * n], Zs);
    out[i] = (int)__ldg(labels + ((long long)tx * Ys + ty) * Zs + tz);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int elem_size(int dtype) {
    switch (dtype) {
        case SKB_U8: return 1;
        case SKB_I16: case SKB_F16: case SKB_BF16: return 2;
        case SKB_I32: case SKB_F32: return 4;
    }
    return 0;
}

static void fill_crop(AsmParams& P, const int32_t crop[3], const int32_t overlap[3]) {
    const int dims[3] = {P.X, P.Y, P.Z};
    long long cropvol = 1;
    for (int a = 0; a < 3; ++a) {
        P.cs[a] = crop[a] < dims[a] ? crop[a] : dims[a];
        P.ov[a] = overlap[a];
        P.step[a] = P.cs[a] - 2 * P.ov[a];
        int k_last = (dims[a] - 1) / P.step[a];
        P.shifted[a] = (k_last * P.step[a] + P.cs[a] > dims[a]) ? 1 : 0;
        cropvol *= P.cs[a];
    }
    P.fast_ok = (P.N == 1 || cropvol <= (1LL << 24)) ? 1 : 0;
    P.single_crop = (P.cs[0] == P.X && P.cs[1] == P.Y && P.cs[2] == P.Z && !P.ov[0] && !P.ov[1] && !P.ov[2]) ? 1 : 0;
}

template <typename VecT, typename OutT>
static void launch_assemble_t(const AsmParams& P, OutT* out, long long v_begin, long long v_end, cudaStream_t st) {
    // voxels [v_begin, v_end) of the flat index; v_begin is a multiple of 256 (checked by the callers)
    const long long c_begin = v_begin / 256, c_end = P.vec_aligned ? v_end / 256 : c_begin;
    if (c_end > c_begin) {
        long long blocks = (c_end - c_begin + ASM_WARPS - 1) / ASM_WARPS;
        // 4 resident CTAs per SM; warps stride over the chunks.  Measured at 2048x2048x512 (ms per launch):
        // 2 CTAs/SM 4.94, 3: 3.98, 4: 3.59, 5 (48 registers): 3.80 — more streams in flight than this hurt.
        if (blocks > 148 * 4) blocks = 148 * 4;
        const char* mode = getenv("SKB_GATHER_STAGING");  // experiments: "bulk" | "ldgsts" | unset = register prefetch
        const bool eligible = P.fast_ok && !P.dense && P.flat_bits && !P.planar;
        const int smem = ASM_WARPS * StageCfg<VecT>::WARP_BYTES;
        if (eligible && mode && mode[0] == 'b') {
            // > 48 KB of dynamic shared memory is opt-in per function (and per device): cheap, so set it every time
            cudaFuncSetAttribute(assemble_tma_kernel<VecT, OutT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(assemble_tma_kernel<VecT, OutT, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            assemble_tma_kernel<VecT, OutT, true><<<(unsigned)blocks, 32 * ASM_WARPS, smem, st>>>(P, out, (unsigned)c_begin, (unsigned)c_end);
        } else if (eligible && mode && mode[0] == 'l') {
            cudaFuncSetAttribute(assemble_tma_kernel<VecT, OutT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(assemble_tma_kernel<VecT, OutT, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            assemble_tma_kernel<VecT, OutT, false><<<(unsigned)blocks, 32 * ASM_WARPS, smem, st>>>(P, out, (unsigned)c_begin, (unsigned)c_end);
        } else {
            // large whole-volume N = 1 passes over the sparse labels: probe the field's density on the device and launch
            // both instantiations (the one not picked returns at once); everything else takes the sparse one
            AsmParams Q = P;
            const bool probe = eligible && P.N == 1 && P.single_crop && P.halo_lo == nullptr && P.halo_hi == nullptr &&
                               P.z_off == 0 && P.Zl == P.Z && P.density != nullptr && c_end - c_begin >= 65536 &&
                               !getenv("SKB_NO_DENSE");
            if (probe) {
                unsigned* cnt = const_cast<unsigned*>(P.density);
                cudaMemsetAsync(cnt, 0, sizeof(unsigned), st);
                assemble_probe_kernel<VecT><<<ASM_PROBE_GROUPS / 256, 256, 0, st>>>(P, (unsigned)c_begin, (unsigned)c_end, cnt);
                Q.dense_from = ASM_PROBE_GROUPS * 8u / 12u;  // >= ~8 % of the sampled voxels carry a vector
                const char* dw = getenv("SKB_DENSE_WORK");     // measurements only
                Q.dense_work = dw ? atoi(dw) : ASM_DENSE_WORK;
                long long dblocks = (c_end - c_begin + ASM_WARPS - 1) / ASM_WARPS;
                if (dblocks > 148 * 3) dblocks = 148 * 3;
                assemble_kernel<VecT, OutT, true><<<(unsigned)dblocks, 32 * ASM_WARPS, 0, st>>>(Q, out, v_end, (unsigned)c_begin, (unsigned)c_end);
            } else {
                Q.density = nullptr;
            }
            assemble_kernel<VecT, OutT, false><<<(unsigned)blocks, 32 * ASM_WARPS, 0, st>>>(Q, out, v_end, (unsigned)c_begin, (unsigned)c_end);
        }
    }
    const long long first = c_end * 256;
    if (first < v_end) {
        const long long rest = (v_end - first + 255) / 256;
        assemble_tail_kernel<VecT, OutT><<<(unsigned)((rest + ASM_WARPS - 1) / ASM_WARPS), 32 * ASM_WARPS, 0, st>>>(P, out, v_end, first);
    }
}

template <typename VecT>
static void launch_assemble(const AsmParams& P, void* out, int out_dtype, long long v_begin, long long v_end, cudaStream_t st) {
    if (out_dtype == SKB_I32) launch_assemble_t<VecT, int32_t>(P, static_cast<int32_t*>(out), v_begin, v_end, st);
    else launch_assemble_t<VecT, int16_t>(P, static_cast<int16_t*>(out), v_begin, v_end, st);
}

extern "C" int skb_assemble(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z, const float scale[3], int N,
                            double decay, const int32_t crop[3], const int32_t overlap[3], const void* workspace,
                            const void* labels_dense, int label_dtype, void* out, int out_dtype, void* stream) {
    return skb_assemble_range(vec, vec_dtype, X, Y, Z, scale, N, decay, crop, overlap, workspace, labels_dense, label_dtype,
                              out, out_dtype, 0, X * Y * Z, stream);
}

extern "C" int skb_assemble_range(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z, const float scale[3],
                                  int N, double decay, const int32_t crop[3], const int32_t overlap[3],
                                  const void* workspace, const void* labels_dense, int label_dtype, void* out,
                                  int out_dtype, int64_t first_voxel, int64_t n_voxels, void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_assemble");
    if (rc) return rc;
    SKB_REQUIRE(first_voxel >= 0 && n_voxels >= 0 && first_voxel + n_voxels <= X * Y * Z && first_voxel % 256 == 0,
                "skb_assemble_range: [first_voxel, first_voxel+n_voxels) must lie in the volume and start on a multiple of 256");
    if (n_voxels == 0) return SKB_OK;
    SKB_REQUIRE(vec && out && scale && crop && overlap, "skb_assemble: NULL pointer");
    SKB_REQUIRE(workspace || labels_dense, "skb_assemble: need a CCL workspace or a dense label volume");
    SKB_REQUIRE(vec_dtype == SKB_F16 || vec_dtype == SKB_BF16 || vec_dtype == SKB_F32, "skb_assemble: vec dtype");
    SKB_REQUIRE(out_dtype == SKB_I32 || out_dtype == SKB_I16, "skb_assemble: out dtype must be i32 or i16");
    SKB_REQUIRE(N >= 1, "skb_assemble: N must be >= 1");
    SKB_REQUIRE(skb_aligned16(out), "skb_assemble: out must be 16-byte aligned");
    if (labels_dense)
        SKB_REQUIRE(label_dtype == SKB_I16 || label_dtype == SKB_I32 || label_dtype == SKB_U8, "skb_assemble: label dtype");
    const bool any_ov = overlap[0] > 0 || overlap[1] > 0 || overlap[2] > 0;
    const bool all_ov = overlap[0] > 0 && overlap[1] > 0 && overlap[2] > 0;
    SKB_REQUIRE(overlap[0] >= 0 && overlap[1] >= 0 && overlap[2] >= 0 && any_ov == all_ov,
                "skb_assemble: overlap must be all zero or all positive (the reference's destination slicing, eval.py:281-283)");
    const int64_t dims[3] = {X, Y, Z};
    for (int a = 0; a < 3; ++a) {
        int64_t cs = crop[a] < dims[a] ? crop[a] : dims[a];
        SKB_REQUIRE(crop[a] > 0 && cs - 2 * overlap[a] > 0,
                    "skb_assemble: crop must exceed twice the overlap on every axis (the reference's crops() would not terminate)");
    }
    AsmParams P = {};
    P.vec = vec; P.vec_hops = vec;
    P.cstride = X * Y * Z;
    P.X = (int)X; P.Y = (int)Y; P.Z = (int)Z;
    P.Zl = (int)Z; P.z_off = 0;
    P.s[0] = scale[0]; P.s[1] = scale[1]; P.s[2] = scale[2];
    P.N = N; P.decay = decay;
    fill_crop(P, crop, overlap);
    P.vec_aligned = skb_aligned16(vec) && ((P.cstride * elem_size(vec_dtype)) % 16 == 0);
    if (labels_dense) {
        P.dense = labels_dense; P.dense_dtype = label_dtype;
    } else {
        SkbCclLayout L = skb_ccl_layout(X, Y, Z, 1);
        const char* base = static_cast<const char*>(workspace);
        P.bits = reinterpret_cast<const ull*>(base + L.off_bits);
        P.parent = reinterpret_cast<const int*>(base + L.off_parent);
        P.ZW = L.ZW;
        P.flat_bits = (Z % 64 == 0) ? 1 : 0;
        // the density probe's counter: a reserved word of the workspace header (the labelling does not touch it again)
        P.density = reinterpret_cast<const unsigned*>(base + offsetof(SkbCclHeader, reserved));
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long v0 = first_voxel, v1 = first_voxel + n_voxels;
    if (vec_dtype == SKB_F16) launch_assemble<__half>(P, out, out_dtype, v0, v1, st);
    else if (vec_dtype == SKB_BF16) launch_assemble<__nv_bfloat16>(P, out, out_dtype, v0, v1, st);
    else launch_assemble<float>(P, out, out_dtype, v0, v1, st);
    SKB_LAUNCH_CHECK("assemble_kernel");
    return SKB_OK;
}

// 2-D mode (BASELINE config 5): a stack of S independent images.  The reference has no 2-D gather of its own —
// index_skeleton_by_embed asserts 5-D input (skeleton.py:671-673) — so the 2-D gather is its 3-D function applied to
// each slice with Z = 1: out[s,x,y] = labels[s, clamp(rint(x + v0*s0)), clamp(rint(y + v1*s1))] with
// _vec2embed2D's embedding (vector_to_embedding.py:50-76).  vec (S,2,X,Y) f16|bf16|f32, labels = the planar CCL
// workspace of the (S,X,Y) stack (skb_ccl_label_sparse with planar = 1; per-slice numbering) or a dense (S,X,Y) volume.
extern "C" int skb_assemble_planar(const void* vec, int vec_dtype, int64_t S, int64_t X, int64_t Y, const float scale[2],
                                   const void* workspace, const void* labels_dense, int label_dtype, void* out,
                                   int out_dtype, void* stream) {
    int rc = skb_check_volume(S, X, Y, "skb_assemble_planar");
    if (rc) return rc;
    SKB_REQUIRE(vec && out && scale, "skb_assemble_planar: NULL pointer");
    SKB_REQUIRE(workspace || labels_dense, "skb_assemble_planar: need a planar CCL workspace or a dense label stack");
    SKB_REQUIRE(vec_dtype == SKB_F16 || vec_dtype == SKB_BF16 || vec_dtype == SKB_F32, "skb_assemble_planar: vec dtype");
    SKB_REQUIRE(out_dtype == SKB_I32 || out_dtype == SKB_I16, "skb_assemble_planar: out dtype must be i32 or i16");
    SKB_REQUIRE(skb_aligned16(out), "skb_assemble_planar: out must be 16-byte aligned");
    if (labels_dense)
        SKB_REQUIRE(label_dtype == SKB_I16 || label_dtype == SKB_I32 || label_dtype == SKB_U8, "skb_assemble_planar: label dtype");
    AsmParams P = {};
    P.vec = vec; P.vec_hops = vec;
    P.planar = 1; P.plane = X * Y;
    P.cstride = X * Y;
    P.X = (int)S; P.Y = (int)X; P.Z = (int)Y;
    P.Zl = (int)Y; P.z_off = 0;
    P.s[0] = 0.f; P.s[1] = scale[0]; P.s[2] = scale[1];
    P.N = 1; P.decay = 1.0;
    const int32_t crop[3] = {(int32_t)S, (int32_t)X, (int32_t)Y}, ov[3] = {0, 0, 0};
    fill_crop(P, crop, ov);
    P.vec_aligned = skb_aligned16(vec) && (P.plane % 8 == 0) && ((P.plane * elem_size(vec_dtype)) % 16 == 0);
    if (labels_dense) {
        P.dense = labels_dense; P.dense_dtype = label_dtype;
    } else {
        SkbCclLayout L = skb_ccl_layout(S, X, Y, 1);
        const char* base = static_cast<const char*>(workspace);
        P.bits = reinterpret_cast<const ull*>(base + L.off_bits);
        P.parent = reinterpret_cast<const int*>(base + L.off_parent);
        P.ZW = L.ZW;
        P.flat_bits = (Y % 64 == 0) ? 1 : 0;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long V = S * X * Y;
    if (vec_dtype == SKB_F16) launch_assemble<__half>(P, out, out_dtype, 0, V, st);
    else if (vec_dtype == SKB_BF16) launch_assemble<__nv_bfloat16>(P, out, out_dtype, 0, V, st);
    else launch_assemble<float>(P, out, out_dtype, 0, V, st);
    SKB_LAUNCH_CHECK("assemble_kernel (planar)");
    return SKB_OK;
}

// Z-slab form: vec/out hold only z in [z_off, z_off+Zl); label targets that leave the slab are answered from the
// neighbours' boundary planes ingested by skb_shard_ingest_runs (and reported through *status when they lie beyond
// the `label_halo` planes those runs describe); with N > 1 the hops that leave the slab inside their crop read the
// neighbours' vector planes from vec_halo_lo / vec_halo_hi.
extern "C" int skb_assemble_slab_ex(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                                    const float scale[3], int N, double decay, const int32_t crop[3], const int32_t overlap[3],
                                    const void* vec_halo_lo, const void* vec_halo_hi, int64_t vec_halo_planes,
                                    int64_t vec_halo_lo_depth, int64_t vec_halo_hi_depth,
                                    const void* workspace, const uint64_t* halo_lo, const uint64_t* halo_hi,
                                    int64_t label_halo_planes, void* out, int out_dtype, int64_t first_voxel,
                                    int64_t n_voxels, uint32_t* status, void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_assemble_slab");
    if (rc) return rc;
    SKB_REQUIRE(vec && out && scale && workspace, "skb_assemble_slab: NULL pointer");
    SKB_REQUIRE(vec_dtype == SKB_F16 || vec_dtype == SKB_BF16 || vec_dtype == SKB_F32, "skb_assemble_slab: vec dtype");
    SKB_REQUIRE(out_dtype == SKB_I32 || out_dtype == SKB_I16, "skb_assemble_slab: out dtype must be i32 or i16");
    SKB_REQUIRE(Z % 64 == 0 && z_off % 64 == 0 && Zl % 64 == 0 && Zl > 0 && z_off >= 0 && z_off + Zl <= Z,
                "skb_assemble_slab: Z, z_off and Zl must be multiples of 64 with the slab inside the volume");
    SKB_REQUIRE(skb_aligned16(out), "skb_assemble_slab: out must be 16-byte aligned");
    SKB_REQUIRE(N >= 1 && label_halo_planes >= 0 && label_halo_planes <= 64 && vec_halo_planes >= 0,
                "skb_assemble_slab: N >= 1, label halo 0..64 planes, vector halo >= 0 planes");
    const int64_t Vs = X * Y * Zl;
    SKB_REQUIRE(first_voxel >= 0 && n_voxels >= 0 && first_voxel + n_voxels <= Vs && first_voxel % 256 == 0,
                "skb_assemble_slab: [first_voxel, first_voxel+n_voxels) must lie in the slab and start on a multiple of 256");
    if (n_voxels == 0) return SKB_OK;
    const int32_t whole[3] = {(int32_t)X, (int32_t)Y, (int32_t)Z}, none[3] = {0, 0, 0};
    if (!crop) crop = whole;
    if (!overlap) overlap = none;
    const bool any_ov = overlap[0] > 0 || overlap[1] > 0 || overlap[2] > 0;
    const bool all_ov = overlap[0] > 0 && overlap[1] > 0 && overlap[2] > 0;
    SKB_REQUIRE(overlap[0] >= 0 && overlap[1] >= 0 && overlap[2] >= 0 && any_ov == all_ov,
                "skb_assemble_slab: overlap must be all zero or all positive");
    const int64_t dims[3] = {X, Y, Z};
    for (int a = 0; a < 3; ++a) {
        int64_t cs = crop[a] < dims[a] ? crop[a] : dims[a];
        SKB_REQUIRE(crop[a] > 0 && cs - 2 * overlap[a] > 0, "skb_assemble_slab: crop must exceed twice the overlap on every axis");
    }
    if (N > 1) {
        // a hop stays inside the owner crop of its voxel: at most crop - overlap - 1 planes from it (the whole crop when
        // there is no overlap), clipped by the volume
        const int64_t cz = crop[2] < Z ? crop[2] : Z;
        const int64_t reach = any_ov ? cz - overlap[2] - 1 : cz - 1;
        const int64_t need_lo = reach < z_off ? reach : z_off, need_hi = reach < Z - z_off - Zl ? reach : Z - z_off - Zl;
        SKB_REQUIRE((need_lo == 0 || (vec_halo_lo && vec_halo_planes >= need_lo)) &&
                        (need_hi == 0 || (vec_halo_hi && vec_halo_planes >= need_hi)),
                    "skb_assemble_slab: N > 1 needs vector halos of crop_z - overlap_z - 1 planes (clipped by the volume) on both faces");
    }
    AsmParams P = {};
    P.vec = vec; P.vec_hops = vec;
    P.cstride = Vs;
    P.X = (int)X; P.Y = (int)Y; P.Z = (int)Z;
    P.Zl = (int)Zl; P.z_off = (int)z_off;
    P.halo_lo = reinterpret_cast<const ull*>(halo_lo);
    P.halo_hi = reinterpret_cast<const ull*>(halo_hi);
    P.label_halo = (int)label_halo_planes;
    P.status = status;
    P.vhalo_lo = vec_halo_lo; P.vhalo_hi = vec_halo_hi; P.vh = (int)vec_halo_planes;
    P.vdepth_lo = (int)(vec_halo_lo_depth > 0 ? vec_halo_lo_depth : vec_halo_planes);
    P.vdepth_hi = (int)(vec_halo_hi_depth > 0 ? vec_halo_hi_depth : vec_halo_planes);
    SKB_REQUIRE(P.vdepth_lo >= P.vh && P.vdepth_hi >= P.vh, "skb_assemble_slab: a vector halo array is shallower than the halo");
    P.s[0] = scale[0]; P.s[1] = scale[1]; P.s[2] = scale[2];
    P.N = N; P.decay = decay;
    fill_crop(P, crop, overlap);
    P.vec_aligned = skb_aligned16(vec) && ((P.cstride * elem_size(vec_dtype)) % 16 == 0);
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, 1);
    const char* base = static_cast<const char*>(workspace);
    P.bits = reinterpret_cast<const ull*>(base + L.off_bits);
    P.parent = reinterpret_cast<const int*>(base + L.off_parent);
    P.ZW = (int)(Zl / 64);
    P.flat_bits = 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long v0 = first_voxel, v1 = first_voxel + n_voxels;
    if (vec_dtype == SKB_F16) launch_assemble<__half>(P, out, out_dtype, v0, v1, st);
    else if (vec_dtype == SKB_BF16) launch_assemble<__nv_bfloat16>(P, out, out_dtype, v0, v1, st);
    else launch_assemble<float>(P, out, out_dtype, v0, v1, st);
    SKB_LAUNCH_CHECK("assemble_kernel (slab)");
    return SKB_OK;
}

// the N = 1, whole-volume-as-one-crop, whole-slab form (kept for callers of the first ABI revision; no range check)
extern "C" int skb_assemble_slab(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                                 const float scale[3], const void* workspace, const uint64_t* halo_lo,
                                 const uint64_t* halo_hi, void* out, int out_dtype, void* stream) {
    return skb_assemble_slab_ex(vec, vec_dtype, X, Y, Z, z_off, Zl, scale, 1, 1.0, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, workspace,
                                halo_lo, halo_hi, 0, out, out_dtype, 0, X * Y * Zl, nullptr, stream);
}

// ---- split gather: host side --------------------------------------------------------------------
static int split_params(const char* who, AsmParams& P, const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z,
                        int64_t z_off, int64_t Zl, const float scale[3], const void* workspace, const void* out, int out_dtype) {
    int rc = skb_check_volume(X, Y, Z, who);
    if (rc) return rc;
    if (!(vec && out && workspace)) { skb_set_error("%s: NULL pointer", who); return SKB_E_ARG; }
    if (!(vec_dtype == SKB_F16 || vec_dtype == SKB_BF16 || vec_dtype == SKB_F32)) { skb_set_error("%s: vec dtype", who); return SKB_E_ARG; }
    if (!(out_dtype == SKB_I32 || out_dtype == SKB_I16)) { skb_set_error("%s: out dtype must be i32 or i16", who); return SKB_E_ARG; }
    if (Z % 64 != 0 || z_off % 64 != 0 || Zl % 64 != 0 || Zl <= 0 || z_off < 0 || z_off + Zl > Z || (X * Y * Zl) % 256 != 0) {
        skb_set_error("%s: Z, z_off and Zl must be multiples of 64 and the slab a multiple of 256 voxels "
                      "(other shapes take the fused skb_assemble)", who);
        return SKB_E_ARG;
    }
    P = AsmParams();
    P.vec = vec; P.vec_hops = vec;
    P.cstride = X * Y * Zl;
    P.X = (int)X; P.Y = (int)Y; P.Z = (int)Z;
    P.Zl = (int)Zl; P.z_off = (int)z_off;
    if (scale) { P.s[0] = scale[0]; P.s[1] = scale[1]; P.s[2] = scale[2]; }
    P.N = 1; P.decay = 1.0;
    const int32_t crop[3] = {(int32_t)X, (int32_t)Y, (int32_t)Z}, ov[3] = {0, 0, 0};
    fill_crop(P, crop, ov);
    P.vec_aligned = skb_aligned16(vec) && ((P.cstride * elem_size(vec_dtype)) % 16 == 0);
    if (!P.vec_aligned || !skb_aligned16(out)) { skb_set_error("%s: vec channels and out must be 16-byte aligned", who); return SKB_E_ARG; }
    SkbCclLayout L = skb_ccl_layout(X, Y, Z, 1);
    const char* base = static_cast<const char*>(workspace);
    P.bits = reinterpret_cast<const ull*>(base + L.off_bits);
    P.parent = reinterpret_cast<const int*>(base + L.off_parent);
    P.ZW = (int)(Zl / 64);
    P.flat_bits = 1;
    return SKB_OK;
}

template <typename VecT>
static void launch_split(const AsmParams& P, bool resolve, void* out, int out_dtype, uint32_t* flags, unsigned n_chunks,
                         int ctas_per_sm, cudaStream_t st) {
    const unsigned per_cta = ASM_WARPS * STREAM_CHUNKS_PER_WARP;
    unsigned stream_grid = (n_chunks + per_cta - 1) / per_cta;
    if (ctas_per_sm > 0 && stream_grid > 148u * (unsigned)ctas_per_sm) stream_grid = 148u * (unsigned)ctas_per_sm;
    unsigned resolve_grid = ((n_chunks + 31u) / 32u + ASM_WARPS - 1) / ASM_WARPS;
    if (resolve_grid > 148u * 8u) resolve_grid = 148u * 8u;
    if (out_dtype == SKB_I32) {
        if (resolve) assemble_resolve_kernel<VecT, int32_t><<<resolve_grid, 32 * ASM_WARPS, 0, st>>>(P, static_cast<int32_t*>(out), flags, n_chunks);
        else assemble_stream_kernel<VecT, int32_t><<<stream_grid, 32 * ASM_WARPS, 0, st>>>(P, static_cast<int32_t*>(out), flags, n_chunks);
    } else {
        if (resolve) assemble_resolve_kernel<VecT, int16_t><<<resolve_grid, 32 * ASM_WARPS, 0, st>>>(P, static_cast<int16_t*>(out), flags, n_chunks);
        else assemble_stream_kernel<VecT, int16_t><<<stream_grid, 32 * ASM_WARPS, 0, st>>>(P, static_cast<int16_t*>(out), flags, n_chunks);
    }
}

static void launch_split_any(const AsmParams& P, int vec_dtype, bool resolve, void* out, int out_dtype, uint32_t* flags,
                             int ctas_per_sm, cudaStream_t st) {
    const unsigned n_chunks = (unsigned)(((long long)P.X * P.Y * P.Zl) / 256);
    if (vec_dtype == SKB_F16) launch_split<__half>(P, resolve, out, out_dtype, flags, n_chunks, ctas_per_sm, st);
    else if (vec_dtype == SKB_BF16) launch_split<__nv_bfloat16>(P, resolve, out, out_dtype, flags, n_chunks, ctas_per_sm, st);
    else launch_split<float>(P, resolve, out, out_dtype, flags, n_chunks, ctas_per_sm, st);
}

extern "C" int skb_assemble_stream(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                                   const void* workspace, uint32_t* group_flags, void* out, int out_dtype, int ctas_per_sm,
                                   void* stream) {
    AsmParams P;
    int rc = split_params("skb_assemble_stream", P, vec, vec_dtype, X, Y, Z, z_off, Zl, nullptr, workspace, out, out_dtype);
    if (rc) return rc;
    SKB_REQUIRE(group_flags && ctas_per_sm >= 0 && ctas_per_sm <= 8, "skb_assemble_stream: NULL group_flags or ctas_per_sm outside 0..8");
    launch_split_any(P, vec_dtype, false, out, out_dtype, group_flags, ctas_per_sm, static_cast<cudaStream_t>(stream));
    SKB_LAUNCH_CHECK("assemble_stream_kernel");
    return SKB_OK;
}

extern "C" int skb_assemble_resolve(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                                    const float scale[3], const void* workspace, const uint64_t* halo_lo,
                                    const uint64_t* halo_hi, const uint32_t* group_flags, void* out, int out_dtype,
                                    void* stream) {
    AsmParams P;
    SKB_REQUIRE(scale, "skb_assemble_resolve: NULL scale");
    int rc = split_params("skb_assemble_resolve", P, vec, vec_dtype, X, Y, Z, z_off, Zl, scale, workspace, out, out_dtype);
    if (rc) return rc;
    SKB_REQUIRE(group_flags, "skb_assemble_resolve: NULL group_flags");
    P.halo_lo = reinterpret_cast<const ull*>(halo_lo);
    P.halo_hi = reinterpret_cast<const ull*>(halo_hi);
    launch_split_any(P, vec_dtype, true, out, out_dtype, const_cast<uint32_t*>(group_flags), 0, static_cast<cudaStream_t>(stream));
    SKB_LAUNCH_CHECK("assemble_resolve_kernel");
    return SKB_OK;
}

extern "C" int skb_vec_embed3d(const void* vec, int vec_dtype, int64_t B, int64_t X, int64_t Y, int64_t Z,
                               const float scale[3], int N, double decay, float* out, void* stream) {
    int rc = skb_check_volume(X, Y, Z, "skb_vec_embed3d");
    if (rc) return rc;
    SKB_REQUIRE(vec && out && scale && B >= 1 && N >= 1, "skb_vec_embed3d: bad argument");
    SKB_REQUIRE(vec_dtype == SKB_F16 || vec_dtype == SKB_BF16 || vec_dtype == SKB_F32, "skb_vec_embed3d: vec dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long V = X * Y * Z;
    const int32_t crop[3] = {(int32_t)X, (int32_t)Y, (int32_t)Z}, ov[3] = {0, 0, 0};
    const int es = elem_size(vec_dtype);
    if (N == 1 && V % 8 == 0 && skb_aligned16(vec) && skb_aligned16(out) && B * V / 8 < (1LL << 31) * 256) {
        if (vec_dtype == SKB_F16) launch_vec_embed_n1_vec8<__half, 3>(vec, out, B, V, (unsigned)Z, (unsigned)Y, scale, st);
        else if (vec_dtype == SKB_BF16) launch_vec_embed_n1_vec8<__nv_bfloat16, 3>(vec, out, B, V, (unsigned)Z, (unsigned)Y, scale, st);
        else launch_vec_embed_n1_vec8<float, 3>(vec, out, B, V, (unsigned)Z, (unsigned)Y, scale, st);
        SKB_LAUNCH_CHECK("vec_embed_n1_vec8_kernel");
        return SKB_OK;
    }
    for (int64_t b = 0; b < B; ++b) {
        AsmParams P = {};
        P.vec = static_cast<const char*>(vec) + (size_t)b * 3 * V * es;
        P.vec_hops = vec;  // reference take() flattens the batch: hops always read batch 0
        P.cstride = V;
        P.X = (int)X; P.Y = (int)Y; P.Z = (int)Z;
        P.Zl = (int)Z; P.z_off = 0;
        P.s[0] = scale[0]; P.s[1] = scale[1]; P.s[2] = scale[2];
        P.N = N; P.decay = decay;
        fill_crop(P, crop, ov);
        float* o = out + (size_t)b * 3 * V;
        unsigned nb = (unsigned)((V + 255) / 256);
        if (vec_dtype == SKB_F16) vec_embed3d_kernel<__half><<<nb, 256, 0, st>>>(P, o, V);
        else if (vec_dtype == SKB_BF16) vec_embed3d_kernel<__nv_bfloat16><<<nb, 256, 0, st>>>(P, o, V);
        else vec_embed3d_kernel<float><<<nb, 256, 0, st>>>(P, o, V);
    }
    SKB_LAUNCH_CHECK("vec_embed3d_kernel");
    return SKB_OK;
}

extern "C" int skb_vec_embed2d(const void* vec, int vec_dtype, int64_t B, int64_t X, int64_t Y, const float scale[2],
                               float* out, void* stream) {
    SKB_REQUIRE(vec && out && scale && B >= 1 && X >= 1 && Y >= 1, "skb_vec_embed2d: bad argument");
    SKB_REQUIRE(X < (1 << 24) && Y < (1 << 24), "skb_vec_embed2d: axis too long for exact fp32 indices");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = B * X * Y;
    unsigned nb = (unsigned)((total + 255) / 256);
    if ((X * Y) % 8 == 0 && skb_aligned16(vec) && skb_aligned16(out) &&
        (vec_dtype == SKB_F16 || vec_dtype == SKB_BF16 || vec_dtype == SKB_F32)) {
        if (vec_dtype == SKB_F16) launch_vec_embed_n1_vec8<__half, 2>(vec, out, B, X * Y, (unsigned)Y, 1u, scale, st);
        else if (vec_dtype == SKB_BF16) launch_vec_embed_n1_vec8<__nv_bfloat16, 2>(vec, out, B, X * Y, (unsigned)Y, 1u, scale, st);
        else launch_vec_embed_n1_vec8<float, 2>(vec, out, B, X * Y, (unsigned)Y, 1u, scale, st);
        SKB_LAUNCH_CHECK("vec_embed_n1_vec8_kernel (2-D)");
        return SKB_OK;
    }
    if (vec_dtype == SKB_F16)
        vec_embed2d_kernel<__half><<<nb, 256, 0, st>>>(static_cast<const __half*>(vec), (int)X, (int)Y, scale[0], scale[1], out, total);
    else if (vec_dtype == SKB_BF16)
        vec_embed2d_kernel<__nv_bfloat16><<<nb, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(vec), (int)X, (int)Y, scale[0], scale[1], out, total);
    else if (vec_dtype == SKB_F32)
        vec_embed2d_kernel<float><<<nb, 256, 0, st>>>(static_cast<const float*>(vec), (int)X, (int)Y, scale[0], scale[1], out, total);
    else SKB_REQUIRE(false, "skb_vec_embed2d: vec dtype");
    SKB_LAUNCH_CHECK("vec_embed2d_kernel");
    return SKB_OK;
}

extern "C" int skb_vec_embed_bwd(const float* grad_out, int64_t B, int C, int64_t inner, const float* scale,
                                 void* grad_vec, int vec_dtype, void* stream) {
    SKB_REQUIRE(grad_out && grad_vec && scale && B >= 1 && (C == 2 || C == 3) && inner >= 1, "skb_vec_embed_bwd: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = B * C * inner;
    unsigned nb = (unsigned)((total + 255) / 256);
    const float s2 = C == 3 ? scale[2] : 0.f;
    if (vec_dtype == SKB_F16)
        vec_embed_bwd_kernel<__half><<<nb, 256, 0, st>>>(grad_out, static_cast<__half*>(grad_vec), inner, C, scale[0], scale[1], s2, total);
    else if (vec_dtype == SKB_BF16)
        vec_embed_bwd_kernel<__nv_bfloat16><<<nb, 256, 0, st>>>(grad_out, static_cast<__nv_bfloat16*>(grad_vec), inner, C, scale[0], scale[1], s2, total);
    else if (vec_dtype == SKB_F32)
        vec_embed_bwd_kernel<float><<<nb, 256, 0, st>>>(grad_out, static_cast<float*>(grad_vec), inner, C, scale[0], scale[1], s2, total);
    else SKB_REQUIRE(false, "skb_vec_embed_bwd: vec dtype");
    SKB_LAUNCH_CHECK("vec_embed_bwd_kernel");
    return SKB_OK;
}

extern "C" int skb_index_by_embed(const void* labels, int label_dtype, int64_t Xs, int64_t Ys, int64_t Zs,
                                  const float* embed, int64_t n, int32_t* out, void* stream) {
    int rc = skb_check_volume(Xs, Ys, Zs, "skb_index_by_embed");
    if (rc) return rc;
    SKB_REQUIRE(labels && embed && out && n >= 0, "skb_index_by_embed: bad argument");
    if (n == 0) return SKB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned nb = (unsigned)((n + 255) / 256);
    if (label_dtype == SKB_I16)
        index_by_embed_kernel<int16_t><<<nb, 256, 0, st>>>(static_cast<const int16_t*>(labels), (int)Xs, (int)Ys, (int)Zs, embed, n, out);
    else if (label_dtype == SKB_I32)
        index_by_embed_kernel<int32_t><<<nb, 256, 0, st>>>(static_cast<const int32_t*>(labels), (int)Xs, (int)Ys, (int)Zs, embed, n, out);
    else if (label_dtype == SKB_U8)
        index_by_embed_kernel<uint8_t><<<nb, 256, 0, st>>>(static_cast<const uint8_t*>(labels), (int)Xs, (int)Ys, (int)Zs, embed, n, out);
    else SKB_REQUIRE(false, "skb_index_by_embed: label dtype must be u8, i16 or i32");
    SKB_LAUNCH_CHECK("index_by_embed_kernel");
    return SKB_OK;
}
