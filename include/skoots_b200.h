/*
 * skoots_b200 — C-ABI of the B200 (sm_100a) skeleton-embedding instance-assembly path.
 *
 * The reference (buswinka/skoots v0.0.5) is pure Python and has no FFI; its "operator API" for
 * this path is the set of `skoots.lib.*` callables its pipelines bind by name.
 * Each entry point below replaces the arithmetic of one of those callables; the Python mirror in
 * `skoots_b200/lib/` keeps their names/signatures and calls these through ctypes
 * (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer owned by the caller (inputs, outputs and workspace);
 *     the library never allocates, frees or synchronises (only exception: the skb_peer_* set-up
 *     calls of section (e')); work is enqueued on `stream`
 *     (a cudaStream_t / CUstream passed as void*; NULL = legacy default stream);
 *   - volumes are C-contiguous with Z fastest: (X,Y,Z), vector fields (3,X,Y,Z) — the
 *     reference's layout (skoots/lib/eval.py:61-64);
 *   - a volume may hold at most 2^31 voxels (voxel indices are non-negative int32), each axis < 2^24;
 *   - return value 0 = enqueued; <0 = SKB_E_* (nothing enqueued), text in skb_last_error();
 *   - asynchronous conditions (workspace overflow) are reported in a caller-provided
 *     device status word, see skb_ccl_*.
 */
#ifndef SKOOTS_B200_H
#define SKOOTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKB_VERSION 201

/* element types */
enum { SKB_U8 = 0, SKB_I16 = 1, SKB_I32 = 2, SKB_F16 = 3, SKB_BF16 = 4, SKB_F32 = 5 };

/* error codes */
enum {
    SKB_OK = 0,
    SKB_E_ARG = -1,       /* bad argument (NULL, size, dtype, alignment) */
    SKB_E_WORKSPACE = -2, /* workspace too small */
    SKB_E_CUDA = -3,      /* CUDA launch error */
    SKB_E_RANGE = -4      /* volume too large for 32-bit voxel indices */
};

/* bits of the device status word written by the CCL kernels */
#define SKB_STATUS_ROOT_OVERFLOW 1u  /* more tile-local components than `capacity` */

int skb_version(void);
const char* skb_last_error(void); /* thread-local, valid until the next failing call */

/* ---------------------------------------------------------------------------------------------
 * a1  vector_to_embedding            skoots/lib/vector_to_embedding.py:135-174 (_vec2embed3D 79-132)
 *   out[b,c,x,y,z] = idx_c + vec*scale_c, then N-1 crop-local hops (round, clamp to [0,dim], fp32
 *   ravel, gather) with decay — bit-exact restatement of the reference's fp32 op sequence.
 *   vec: (B,3,X,Y,Z) f16|bf16|f32, out: (B,3,X,Y,Z) f32.  B>1 with N>1 reproduces the reference's
 *   `take` over the flattened batch (all batches index batch 0).
 * ------------------------------------------------------------------------------------------- */
int skb_vec_embed3d(const void* vec, int vec_dtype, int64_t B, int64_t X, int64_t Y, int64_t Z,
                    const float scale[3], int N, double decay, float* out, void* stream);

/* 2-D form, vector_to_embedding.py:50-76.  vec (B,2,X,Y) -> out (B,2,X,Y) f32 */
int skb_vec_embed2d(const void* vec, int vec_dtype, int64_t B, int64_t X, int64_t Y,
                    const float scale[2], float* out, void* stream);

/* backward of the N=1 forms: grad_vec[b,c,...] = grad_out[b,c,...] * scale_c, cast to vec_dtype.
 * C = 2 or 3 channels of `inner` elements each. */
int skb_vec_embed_bwd(const float* grad_out, int64_t B, int C, int64_t inner, const float* scale,
                      void* grad_vec, int vec_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a2  index_skeleton_by_embed        skoots/lib/skeleton.py:656-695
 *   out[i] = labels[clamp(rint(e_x),0,Xs-1), clamp(rint(e_y),..), clamp(rint(e_z),..)] as int32.
 *   labels (Xs,Ys,Zs) i16|i32|u8 ; embed (3,n) f32 (n = x*y*z of the crop) ; out (n) i32
 * ------------------------------------------------------------------------------------------- */
int skb_index_by_embed(const void* labels, int label_dtype, int64_t Xs, int64_t Ys, int64_t Zs,
                       const float* embed, int64_t n, int32_t* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a3  efficient_flood_fill (connected components of mask>0)   skoots/lib/flood_fill.py:13-140
 *   6-connectivity in 3-D (scipy.ndimage.label default structure, flood_fill.py:135), or, with
 *   `planar` != 0, 4-connectivity inside each x-plane with numbering restarting per plane
 *   (utils/flood_and_stitch.py:63-69 applied to a stack (S,X,Y) passed as (X=S,Y=X,Z=Y)).
 *   Labels are label_base + 1 + (raster-order rank of the component's first voxel): identical to
 *   scipy's numbering; the reference's single-crop result is label_base = 2 (flood_fill.py:138).
 *
 *   The labelling is produced in a sparse form in `workspace` (skb_ccl_label_sparse); it can then
 *   be densified (skb_ccl_write_dense) or consumed directly by the fused gather (skb_assemble).
 *   `capacity` = max number of tile-local components the workspace can hold (worst case V/2+1);
 *   if exceeded, SKB_STATUS_ROOT_OVERFLOW is OR-ed into *status (device) and labels are invalid.
 *   ncomp (device int32, may be NULL) receives the number of components.
 * ------------------------------------------------------------------------------------------- */
size_t skb_ccl_workspace_bytes(int64_t X, int64_t Y, int64_t Z, int64_t capacity);

/* flags: SKB_CCL_WORKSPACE_CLEAN = this workspace was last used by a completed labelling pass of the
 * same volume shape (every pass leaves its root bitmap zeroed), so the V/8-byte memset is skipped. */
#define SKB_CCL_WORKSPACE_CLEAN 1
/* run only part of the labelling (both skb_ccl_label_sparse and skb_shard_label_local), so that a caller
 * can start skb_assemble_stream as soon as the bit mask exists: PHASE_PACK = header, clears and the
 * mask -> bit-mask pack; PHASE_LABEL = everything after it.  Neither bit = the whole labelling. */
#define SKB_CCL_PHASE_PACK 2
#define SKB_CCL_PHASE_LABEL 4
/* how the tile kernel hands tiles to its persistent warps: by default batches are claimed dynamically when
 * there are >= 32 tiles per warp and tiles are assigned statically otherwise; these force one or the other
 * (same labels either way — for tests and measurements) */
#define SKB_CCL_TILES_DYNAMIC 8
#define SKB_CCL_TILES_STATIC 16
/* do not reset *status at the start of the pass: the status word then accumulates over a series of passes (a timed
 * loop, a replayed CUDA graph) and the caller clears it when it reads it */
#define SKB_CCL_KEEP_STATUS 32
int skb_ccl_label_sparse(const void* mask, int mask_dtype, int64_t X, int64_t Y, int64_t Z,
                         int planar, int32_t label_base, int64_t capacity, void* workspace,
                         size_t workspace_bytes, int32_t* ncomp, uint32_t* status, int flags,
                         void* stream);

/* dense labels from the sparse form. out (X,Y,Z) i16|i32, may alias the mask given to
 * skb_ccl_label_sparse (the reference labels in place, flood_fill.py:50). */
int skb_ccl_write_dense(const void* workspace, int64_t X, int64_t Y, int64_t Z, void* out,
                        int out_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a1+a2+a6  fused instance assembly   skoots/lib/eval.py:245-284 (+ cropper.py:58-144)
 *   For every voxel: find the reference crop that owns it (crop/overlap grid, later crops
 *   overwrite, outer `overlap` margin never written -> 0), walk N hops in that crop's local
 *   frame, add the crop origin in fp32 (eval.py:274), round/clamp against the whole volume and
 *   read the component label there.  crop >= dims and overlap 0 = "whole volume is one crop".
 *   Labels come either from the sparse CCL workspace (labels_dense == NULL) or from a dense
 *   label volume (labels_dense != NULL, dtype label_dtype; workspace may then be NULL).
 *   vec (3,X,Y,Z) f16|bf16|f32 ; out (X,Y,Z) i32|i16.
 * ------------------------------------------------------------------------------------------- */
int skb_assemble(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z,
                 const float scale[3], int N, double decay, const int32_t crop[3],
                 const int32_t overlap[3], const void* workspace, const void* labels_dense,
                 int label_dtype, void* out, int out_dtype, void* stream);

/* the same for the voxels [first_voxel, first_voxel + n_voxels) of the flat (X,Y,Z) index only
 * (first_voxel a multiple of 256): lets a caller pipeline host->device uploads of X-slabs of the vector
 * field against the gather and the device->host download of finished slabs. */
int skb_assemble_range(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z,
                       const float scale[3], int N, double decay, const int32_t crop[3],
                       const int32_t overlap[3], const void* workspace, const void* labels_dense,
                       int label_dtype, void* out, int out_dtype, int64_t first_voxel,
                       int64_t n_voxels, void* stream);

/* a10  2-D mode (BASELINE.json configs[4]): a stack of S independent images.  The reference has no 2-D gather of
 *   its own (index_skeleton_by_embed asserts 5-D input, skoots/lib/skeleton.py:671-673): the 2-D gather is its 3-D
 *   function applied per slice with Z = 1, on _vec2embed2D's embedding (vector_to_embedding.py:50-76):
 *   out[s,x,y] = labels[s, clamp(rint(x + v[s,0,x,y]*scale0), 0, X-1), clamp(rint(y + v[s,1,x,y]*scale1), 0, Y-1)].
 *   vec (S,2,X,Y) f16|bf16|f32; labels = the workspace of skb_ccl_label_sparse(planar = 1) over the (S,X,Y) stack
 *   (4-connectivity, numbering restarts per slice — utils/flood_and_stitch.py:63-69) or a dense (S,X,Y) volume;
 *   out (S,X,Y) i32|i16. */
int skb_assemble_planar(const void* vec, int vec_dtype, int64_t S, int64_t X, int64_t Y, const float scale[2],
                        const void* workspace, const void* labels_dense, int label_dtype, void* out,
                        int out_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a4  binary_dilation / binary_dilation_2d / binary_erosion     skoots/lib/morphology.py:130-199
 *   zero-padded 3x3x3 max (op 0), 3x3x1 max (op 1), 3x3x3 min (op 2) over n_volumes = B*C
 *   volumes of (X,Y,Z) fp32.  Out of place.
 * ------------------------------------------------------------------------------------------- */
int skb_stencil3(const float* in, float* out, int64_t n_volumes, int64_t X, int64_t Y, int64_t Z,
                 int op, void* stream);

/* average_baked_skeletons, skoots/lib/skeleton.py:18-48: sum(3x3x3 window)/max(1,count(window>0)) */
int skb_masked_mean27(const float* in, float* out, int64_t n_volumes, int64_t X, int64_t Y,
                      int64_t Z, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a5  tile epilogue                                              skoots/lib/eval.py:145-176
 *   unet: one tile of network output (C>=5, tx,ty,tz) f32|f16|bf16: channels 0..2 vectors,
 *   C-2 skeleton, C-1 probability.  Writes the tile interior (tile minus `overlap` on each side)
 *   at `origin` into vectors_f16 (3,X,Y,Z) and skeleton_u8 (X,Y,Z).
 *   tile/origin/overlap are HOST arrays.
 * ------------------------------------------------------------------------------------------- */
int skb_tile_epilogue(const void* unet, int in_dtype, int C, const int32_t tile[3],
                      const int32_t origin[3], const int32_t overlap[3], float threshold,
                      void* vectors_f16, uint8_t* skeleton_u8, int64_t X, int64_t Y, int64_t Z,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * a7  baked_embed_to_prob                                skoots/lib/embedding_to_prob.py:5-51
 *   embedding (B,C,inner) f32, baked (B,C,inner) f32|f16|bf16, sigma: HOST array of C floats,
 *   out (B,1,inner) f32.  C = 2 or 3.  _bwd writes grad_embedding (f32) and/or grad_baked
 *   (baked dtype); either may be NULL.
 * ------------------------------------------------------------------------------------------- */
int skb_embed_prob_fwd(const float* embedding, const void* baked, int baked_dtype, int64_t B, int C,
                       int64_t inner, const float* sigma, float eps, float* out, void* stream);
int skb_embed_prob_bwd(const float* embedding, const void* baked, int baked_dtype,
                       const float* prob, const float* grad_out, int64_t B, int C, int64_t inner,
                       const float* sigma, float eps, float* grad_embedding, void* grad_baked,
                       void* stream);

/* fused vector_to_embedding(N=1) + baked_embed_to_prob (train/engine.py:465-466): vec (B,C,X,Y[,Z]).
 * grad_out == NULL -> forward, writes prob (B,1,...) ; else backward, writes grad_vec (vec dtype). */
int skb_vec_prob(const void* vec, int vec_dtype, const void* baked, int baked_dtype, int64_t B, int C,
                 int64_t X, int64_t Y, int64_t Z, const float* scale, const float* sigma, float eps,
                 float* prob, const float* grad_out, void* grad_vec, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a8  bake_skeleton (CPU/torch semantics) + average_baked_skeletons     skoots/lib/skeleton.py:370-445, 18-48
 *   One launch for a batch of B samples (the reference calls bake_skeleton once per sample inside the data
 *   loader, dataloader.py:177; B = 1 is that call).  masks (B,X,Y,Z) u8|i16|i32 object ids.  Tables (device):
 *   ids = every sample's SORTED object ids, concatenated; id_begin (B+1) = range of sample b in ids;
 *   offsets (n_ids+1) = prefix of point counts over the whole batch into points_xyzw (n_points x 4 floats,
 *   16-byte rows).  anisotropy: HOST array.  baked (B,3,X,Y,Z) f32 = nearest point of the voxel's own skeleton
 *   (first minimum of the sqrt'ed distances), 0 on background; average != 0 fuses the masked 3x3x3 mean of
 *   average_baked_skeletons (sum of the window / max(1, count of entries > 0), zero padded) so the un-averaged
 *   field never reaches HBM.  distance (B,X,Y,Z) f32 optional.  *status |= SKB_STATUS_MISSING_ID when a mask id
 *   has no skeleton (the reference raises KeyError, skeleton.py:422); the caller zeroes *status and may read it
 *   once per batch.
 *   triton_block (device, B ints) != NULL selects the semantics of the reference's Triton kernel instead
 *   (skeleton.py:51-251, what bake_skeleton dispatches for a CUDA mask, :505-512): entry b = the SKEL_BLOCK_SIZE
 *   of its launch (next power of two of the sample's longest skeleton, :361; 0 = no points: zeros).  Differences
 *   from the CPU path: anisotropy weighs the squared differences; lanes past a skeleton's length act as a point
 *   at the origin; ties take the per-axis maximum; an id without a skeleton gives zeros, no error; baked and
 *   distance are fp16 values (stored here as f32).  Pinned for integer-valued coordinates and anisotropy
 *   (tests/golden/bake_triton.npz, generated by the reference on a B200).
 * ------------------------------------------------------------------------------------------- */
#define SKB_STATUS_MISSING_ID 2u
int skb_bake_skeletons(const void* masks, int mask_dtype, int64_t B, int64_t X, int64_t Y, int64_t Z,
                       const int32_t* ids, const int32_t* id_begin, const int32_t* offsets, int n_ids,
                       const float* points_xyzw, int n_points, const float anisotropy[3], int average,
                       const int32_t* triton_block, float* baked, float* distance, uint32_t* status, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a9  skeleton_to_mask                                         skoots/lib/skeleton.py:531-593
 *   points (n_points,3) f32 device; offsets (n_offsets,3) i32 device (the disk stamp of
 *   lib/utils.py:421-438); out (X,Y,Z) f32 must be zeroed by the caller; sets 1.0 at
 *   trunc(point + offset) where inside the volume.
 * ------------------------------------------------------------------------------------------- */
int skb_stamp_disks(const float* points_xyz, int n_points, const int32_t* offsets_xyz, int n_offsets,
                    int64_t X, int64_t Y, int64_t Z, float* out_zeroed, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (e)  Z-sharded post-processing: one rank per GPU owns the slab z in [z_off, z_off+Zl) of the
 *   mask and of the vector field (Z, z_off, Zl multiples of 64).  New design — the reference has no
 *   collective on this path (skoots/lib/eval.py:223-284 is single-process).  Each rank's union-find
 *   lives in the GLOBAL voxel index space of a full-size workspace, so component ids agree across
 *   ranks.  Call order per rank (the two exchanges are NCCL send/recv and all-gather, driven by the
 *   caller — skoots_b200/sharded.py):
 *     skb_shard_clear_halo (per face; the halo words start zeroed and are kept clean this way) ->
 *     skb_shard_label_local -> skb_shard_emit_runs (low / high H planes) -> [send/recv with the
 *     Z-neighbours] -> skb_shard_boundary_pairs (halo_hi = NULL: zeroes the counters of
 *     [n_roots, n_pairs, roots[cap_roots], pairs[2*cap_pairs]] and packs my roots) ->
 *     skb_shard_ingest_runs (into zeroed per-row halo words; the upper neighbour's call also appends the
 *     (my root, its root) pairs of my last plane) -> [all-gather] -> skb_shard_merge -> skb_assemble_slab.
 *     (skb_shard_boundary_pairs with halo_hi != NULL, called AFTER the ingests, finds the pairs by scanning
 *     the halo words instead — the older, slower order; kept for callers that ingest without `exchange`.)
 *   Numbering after the merge is the single-GPU numbering (label_base + 1 + raster rank).
 *   runs buffers hold 3*(cap+1) int32: [count,_,_] then (start voxel, length, root id) triples.
 * ------------------------------------------------------------------------------------------- */
int skb_shard_label_local(const void* mask, int mask_dtype, int64_t X, int64_t Y, int64_t Z,
                          int64_t z_off, int64_t Zl, int64_t capacity, void* workspace,
                          size_t workspace_bytes, uint32_t* status, int flags, void* stream);
/* face_is_high: 0 = the first `halo` planes of the slab, 1 = its last (halo <= 64) */
int skb_shard_emit_runs(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                        int face_is_high, int64_t halo, int32_t* runs, int64_t cap, uint32_t* status,
                        void* stream);
/* exchange != NULL (the UPPER neighbour's runs): also appends the (my root, neighbour root) pairs of runs that touch
 * my last plane to the exchange buffer — call skb_shard_boundary_pairs(halo_hi = NULL) BEFORE it, which zeroes the
 * buffer's counters and packs my roots */
int skb_shard_ingest_runs(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                          const int32_t* runs, int64_t cap, uint64_t* halo_words_zeroed, int32_t* exchange,
                          int64_t cap_roots, int64_t cap_pairs, uint32_t* status, void* stream);
/* instead of re-zeroing all X*Y halo words before every ingest: zero exactly the words the previous pass's ingest
 * set, from that pass's run list (prev_runs = the receive buffer, before the next exchange overwrites it) */
int skb_shard_clear_halo(int64_t Z, const int32_t* prev_runs, int64_t cap, uint64_t* halo_words, void* stream);
int skb_shard_boundary_pairs(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t z_off,
                             int64_t Zl, int64_t capacity, const uint64_t* halo_hi, int32_t* exchange,
                             int64_t cap_roots, int64_t cap_pairs, uint32_t* status, void* stream);
int skb_shard_merge(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t capacity,
                    const int32_t* gathered, int world, int rank, int64_t cap_roots, int64_t cap_pairs,
                    int32_t label_base, int32_t* ncomp, uint32_t* status, void* stream);
/* fused gather on a slab (N = 1, whole volume as one crop): vec (3,X,Y,Zl), out (X,Y,Zl);
 * halo_lo / halo_hi (X*Y words each, NULL at the volume's ends) come from skb_shard_ingest_runs */
int skb_assemble_slab(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z, int64_t z_off,
                      int64_t Zl, const float scale[3], const void* workspace, const uint64_t* halo_lo,
                      const uint64_t* halo_hi, void* out, int out_dtype, void* stream);

/* The general slab gather.  On top of skb_assemble_slab:
 *   - N hops with `decay` over the reference's crop grid (crop / overlap as in skb_assemble; NULL = the whole volume
 *     as one crop).  A hop stays inside the owner crop of its voxel, so it can leave the slab by at most
 *     crop_z - overlap_z - 1 planes: vec_halo_lo / vec_halo_hi hold the vector field's planes
 *     [z_off - vec_halo_planes, z_off) and [z_off+Zl, z_off+Zl+vec_halo_planes) of the field's dtype (NULL at the
 *     volume's ends; not needed for N = 1), as the LAST / FIRST vec_halo_planes planes of (3,X,Y,depth) arrays:
 *     vec_halo_*_depth = 0 or vec_halo_planes -> packed copies of the neighbours' faces; = the neighbour's slab depth ->
 *     the neighbour's OWN slab, mapped into this process (CUDA IPC, skb_peer_*): the few hops that cross a face then read
 *     the peer GPU's memory over NVLink directly and nothing is exchanged;
 *   - label_halo_planes = how many planes beyond each face the halo words describe (the `halo` given to
 *     skb_shard_emit_runs*).  A gather target beyond them cannot be answered from this rank's data:
 *     SKB_STATUS_HALO_RANGE is OR-ed into *status instead of returning a wrong label silently.  0 = do not check;
 *   - only the voxels [first_voxel, first_voxel + n_voxels) of the slab's flat (X,Y,Zl) index are written
 *     (first_voxel a multiple of 256), so a caller can pipeline X-slabs of the upload / gather / download. */
#define SKB_STATUS_HALO_RANGE 16u
int skb_assemble_slab_ex(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                         const float scale[3], int N, double decay, const int32_t crop[3], const int32_t overlap[3],
                         const void* vec_halo_lo, const void* vec_halo_hi, int64_t vec_halo_planes,
                         int64_t vec_halo_lo_depth, int64_t vec_halo_hi_depth,
                         const void* workspace, const uint64_t* halo_lo, const uint64_t* halo_hi,
                         int64_t label_halo_planes, void* out, int out_dtype, int64_t first_voxel,
                         int64_t n_voxels, uint32_t* status, void* stream);

/* Split form of the fused gather for N = 1 with the whole volume as one crop (the headline mode), on a
 * volume (z_off = 0, Zl = Z) or a slab.  skb_assemble_stream needs only the bit mask of the CCL
 * workspace (i.e. the SKB_CCL_PHASE_PACK part of the labelling): it streams the vector field, stores
 * the zeros of every 8-voxel group that cannot have a label and records the other groups in
 * group_flags (one uint32 per 256 voxels, X*Y*Zl/256 words).  The caller runs it on a second stream
 * next to the rest of the labelling (SKB_CCL_PHASE_LABEL), then skb_assemble_resolve fills in the
 * flagged groups.  Together they write exactly what skb_assemble / skb_assemble_slab write.
 * Z, z_off, Zl multiples of 64 and X*Y*Zl a multiple of 256; other shapes use the fused call.
 * ctas_per_sm: 0 = one CTA per 32 chunks; 1..8 = a persistent grid of that many 256-thread CTAs per SM,
 * which bounds the slice of every SM the stream phase holds while the labelling kernels run next to it. */
int skb_assemble_stream(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z, int64_t z_off,
                        int64_t Zl, const void* workspace, uint32_t* group_flags, void* out,
                        int out_dtype, int ctas_per_sm, void* stream);
int skb_assemble_resolve(const void* vec, int vec_dtype, int64_t X, int64_t Y, int64_t Z, int64_t z_off,
                         int64_t Zl, const float scale[3], const void* workspace,
                         const uint64_t* halo_lo, const uint64_t* halo_hi, const uint32_t* group_flags,
                         void* out, int out_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (e')  The same pass with both exchanges done by the kernels themselves over NVLink peer memory —
 *   no NCCL call, no host involvement between the phases, the whole pass is one CUDA graph.
 *   Every rank owns a MAILBOX in peer-visible memory (layout identical on all ranks; two copies of
 *   every receive buffer, pass k uses copy k & 1).  Producers store straight into the consumer's
 *   mailbox and then release a flag (= the pass number) there; consumers spin on flags in their own
 *   HBM.  A flag that does not arrive within ~8 s sets SKB_STATUS_PEER_TIMEOUT instead of hanging.
 *   Call order per rank and pass (skoots_b200/sharded.py, transport "peer"):
 *     skb_shard_begin_pass (pass counter, halo words of the previous pass cleared, pair counter zeroed)
 *     -> skb_shard_label_local -> skb_shard_emit_runs_peer (both faces + a one-warp signal kernel)
 *     -> skb_shard_ingest_runs_peer (from the low / high neighbour; the latter appends the face pairs)
 *     -> skb_shard_push (roots + pairs to every rank; its last CTA signals) -> skb_shard_merge_peer (eight launches; a
 *     one-kernel cooperative form exists behind SKB_SHARD_FUSED=1 — measured slower) -> skb_assemble_slab[_ex].
 *     21 kernel launches + 1 memset per pass on a rank with two neighbours (round 1: 24 + 3 memsets).
 *   The skb_peer_* calls are set-up / tear-down only: they are the one place the library allocates
 *   (cudaMalloc: legacy CUDA IPC cannot export a caching allocator's sub-allocations).
 * ------------------------------------------------------------------------------------------- */
#define SKB_PEER_HANDLE_BYTES 64
#define SKB_MAX_WORLD 16
#define SKB_STATUS_PEER_TIMEOUT 4u
int skb_peer_alloc(size_t bytes, void** ptr); /* zero-filled device memory on the current device */
int skb_peer_free(void* ptr);
int skb_peer_export(void* ptr, uint8_t handle[SKB_PEER_HANDLE_BYTES]);
int skb_peer_open(const uint8_t handle[SKB_PEER_HANDLE_BYTES], void** ptr); /* maps + enables peer access */
int skb_peer_close(void* ptr);

size_t skb_shard_mailbox_bytes(int world, int64_t cap_runs, int64_t cap_roots, int64_t cap_pairs);
/* starts pass k+1: bumps the mailbox's pass counter, clears its run counters */
int skb_shard_begin(void* mailbox, int world, int64_t cap_runs, int64_t cap_roots, int64_t cap_pairs,
                    void* stream);
/* the same plus, in the same call: skb_shard_clear_halo_peer for both faces with ONE launch (halo_lo / halo_hi: the
 * X*Y halo words of each face, NULL where there is no neighbour) and the zeroing of the exchange buffer's counters
 * (exchange may be NULL) — the prologue of a pass as two launches instead of five nodes */
int skb_shard_begin_pass(void* mailbox, int world, int64_t cap_runs, int64_t cap_roots, int64_t cap_pairs, int64_t Z,
                         uint64_t* halo_lo, uint64_t* halo_hi, int32_t* exchange, void* stream);
/* skb_shard_clear_halo on the mailbox's receive-buffer copy of the PREVIOUS pass; call after skb_shard_begin */
int skb_shard_clear_halo_peer(int64_t Z, void* mailbox, int from_high, int world, int64_t cap_runs,
                              int64_t cap_roots, int64_t cap_pairs, uint64_t* halo_words, void* stream);
/* like skb_shard_emit_runs for BOTH faces of the slab with one launch (`halo` planes each): the triples
 * are stored into the neighbours' mailboxes (NULL = no neighbour on that side) and their flags released */
int skb_shard_emit_runs_peer(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                             int64_t halo, void* mailbox, void* lo_neighbour_mailbox,
                             void* hi_neighbour_mailbox, int world, int64_t cap_runs, int64_t cap_roots,
                             int64_t cap_pairs, uint32_t* status, void* stream);
/* waits for the neighbour's flag, then skb_shard_ingest_runs on the mailbox's receive buffer */
int skb_shard_ingest_runs_peer(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t z_off, int64_t Zl,
                               void* mailbox, int from_high, int world, int64_t cap_runs, int64_t cap_roots,
                               int64_t cap_pairs, uint64_t* halo_words_zeroed, int32_t* exchange,
                               uint32_t* status, void* stream);
/* the all-gather: stores my payload — the slab's root list straight from the workspace, the face pairs the upper
 * neighbour's ingest appended to `exchange` (counters zeroed by skb_shard_begin_pass) — into slot `rank` of every
 * rank's mailbox (peer_mailboxes: HOST array of `world` device pointers, own included); the last CTA of the kernel
 * releases the flags.  No skb_shard_boundary_pairs call is needed on this transport. */
int skb_shard_push(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t capacity, const int32_t* exchange, void* mailbox,
                   const uint64_t* peer_mailboxes, int world, int rank, int64_t cap_runs, int64_t cap_roots,
                   int64_t cap_pairs, uint32_t* status, void* stream);
/* waits for every rank's flag, then skb_shard_merge on the mailbox's gather buffer */
int skb_shard_merge_peer(void* workspace, int64_t X, int64_t Y, int64_t Z, int64_t capacity,
                         void* mailbox, int world, int rank, int64_t cap_runs, int64_t cap_roots,
                         int64_t cap_pairs, int32_t label_base, int32_t* ncomp, uint32_t* status,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * (f2)  after the assembly: renumbering and the validation metrics
 *   Label values must be non-negative and < table_size (instance labels are 0..N+2 on this path); a label
 *   outside the table sets SKB_STATUS_LABEL_RANGE in *status and is left alone / ignored.
 *
 *   skb_renumber     fastremap.renumber(mask, in_place=True), skoots/lib/eval.py:304 — labels become 1..N in
 *                    order of first appearance in the C-order scan, 0 stays 0.  In place; remap (table_size
 *                    int32) receives old label -> new label (0 = absent), *n_labels the count.
 *   skb_unique_index sorted unique labels > 0 (gt.unique(), validate/lib.py:201-205): index[l] = rank of l
 *                    among the labels present, or -1; values (optional, >= count ints) = the labels in order.
 *   skb_contingency  one pass over (gt, pred): inter[i*M+j] = |gt==a_i & pred==b_j|, areas — every
 *                    `logical_and(_a,_b).sum()`, `_a.sum()` of validate/lib.py:211-226 and 253-273.
 *   skb_iou_dice     iou = I/(A+B-I), dice = 2I/(A+B) as ONE fp32 division of the counts converted to fp32
 *                    (the reference divides 0-dim int64 tensors), 0 where the objects do not touch.
 *   skb_accuracies_from_iou   validate/lib.py:170-187: out3 = [true positives, false positives, false
 *                    negatives] at threshold thr; hits_scratch holds N+M int32.
 * ------------------------------------------------------------------------------------------- */
#define SKB_STATUS_LABEL_RANGE 8u
/* largest label (>= 0) of an i16 | i32 volume -> *max_out (device): sizes the tables below */
int skb_label_max(const void* labels, int dtype, int64_t n_voxels, int32_t* max_out, void* stream);
size_t skb_renumber_workspace_bytes(int64_t n_voxels, int64_t table_size);
int skb_renumber(void* labels, int dtype, int64_t n_voxels, int64_t table_size, void* workspace,
                 size_t workspace_bytes, int32_t* remap, int32_t* n_labels, uint32_t* status, void* stream);
/* labels[i] = table[labels[i]] for 0 < labels[i] < table_size, in place: the `replace` step of the reference's
 * multi-crop efficient_flood_fill (flood_fill.py:206-234) as one pass (row f3) */
int skb_apply_label_table(void* labels, int dtype, int64_t n_voxels, const int32_t* table, int64_t table_size,
                          void* stream);
size_t skb_unique_index_workspace_bytes(int64_t table_size);
int skb_unique_index(const void* labels, int dtype, int64_t n_voxels, int64_t table_size, int32_t* index,
                     int32_t* values, int32_t* count, void* workspace, size_t workspace_bytes,
                     uint32_t* status, void* stream);
int skb_contingency(const void* gt, int gt_dtype, const void* pred, int pred_dtype, int64_t n_voxels,
                    const int32_t* index_gt, int64_t table_gt, const int32_t* index_pred, int64_t table_pred,
                    int64_t N, int64_t M, int32_t* inter_zeroed, int32_t* area_gt_zeroed,
                    int32_t* area_pred_zeroed, void* stream);
int skb_iou_dice(const int32_t* inter, const int32_t* area_gt, const int32_t* area_pred, int64_t N, int64_t M,
                 float* iou, float* dice, void* stream);
int skb_accuracies_from_iou(const float* iou, int64_t N, int64_t M, float thr, int32_t* hits_scratch,
                            int32_t* out3, void* stream);

/* ---------------------------------------------------------------------------------------------
 * (f4)  the elastic deformation of the training augmentation     skoots/train/merged_transform.py:75-188, 43-72
 *   noise (3,D,H,W) f32 device = the coarse random field the reference draws with torch.rand (D along X, H along Y,
 *   W along Z); magnitude_rev = displacement_magnitude reversed (HOST array).  Neither of the reference's two dense
 *   (X,Y,Z,3) grids is built: the trilinear displacement is evaluated per voxel / per point from the coarse field.
 *   skb_elastic_resample  F.grid_sample(mode="nearest", align_corners=True, zeros padding) of n_volumes volumes
 *                         (X,Y,Z) f32 through grid = identity + displacement.  Out of place.
 *   skb_elastic_points    new position of every skeleton point (n,3), int64 (points_are_int64 != 0: results truncated
 *                         toward zero like the reference's in-place assignment) or f32; points outside the volume are
 *                         copied unchanged.
 *   Parity is stated with a tolerance (ATen's own CPU and CUDA upsampling kernels differ in the last bit): see DESIGN.md.
 * ------------------------------------------------------------------------------------------- */
int skb_elastic_resample(const float* noise, int64_t D, int64_t H, int64_t W, const float magnitude_rev[3],
                         const float* in, float* out, int64_t n_volumes, int64_t X, int64_t Y, int64_t Z,
                         void* stream);
int skb_elastic_points(const float* noise, int64_t D, int64_t H, int64_t W, const float magnitude_rev[3],
                       const void* points, int points_are_int64, int64_t n_points, int64_t X, int64_t Y,
                       int64_t Z, void* out_points, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SKOOTS_B200_H */
