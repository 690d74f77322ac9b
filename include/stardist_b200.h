/* stardist_b200.h -- C ABI of libstardist_b200.so (B200 / sm_100a prediction hot path).
 *
 * Two families of entry points:
 *
 *  (1) _LIB_*  : HOST pointers, blocking.  These are the drop-in boundary: the first two have
 *      exactly the signatures of the reference's own C ABI
 *      (stardist/lib/stardist3d_lib.h:55-82, implemented by stardist3d_lib.c:5-41 on top of
 *      _COMMON_* in stardist3d_impl.cpp:956,1404); the 2D ones are the analogues of the CPython
 *      entry point c_non_max_suppression_inds (stardist/lib/stardist2d.cpp:390-615, called from
 *      stardist/nms.py:220-225) and of polygons_to_label (stardist/geometry/geom2d.py:149-197).
 *      Inputs are copied to the device, the CUDA kernels run, the result is copied back.
 *
 *  (2) sdb_*   : DEVICE pointers + cudaStream_t, stream ordered.  Used by the Python host layer
 *      (stardist_b200/) to keep the whole predict_instances pipeline resident in HBM.
 *
 * All functions returning int return 0 on success, non-zero on error; sdb_last_error() gives the
 * message.  There is no CPU fallback: without a usable CUDA device every call fails.
 */
#ifndef STARDIST_B200_H
#define STARDIST_B200_H

#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
#include <cstdbool>
extern "C" {
#else
#include <stdbool.h>
#endif

typedef void* sdb_stream_t; /* a cudaStream_t (0 = default stream) */

/* ----------------------------------------------------------------------------- misc */
const char* sdb_last_error(void);
int sdb_device_info(int* n_devices, int* sm_count, int* cc_major, int* cc_minor);
/* number of kernel launches issued by this library since the last reset (bench "gpu_launches") */
long long sdb_launch_count(int reset);

/* optional live profiling with CUDA events on the launching stream (bench.py's roofline): names are
 * "nms2d_clip" (units = polygon pairs tested) and "conv_tc" (units = algorithmic FLOPs) */
int sdb_profile_enable(int on);
int sdb_profile_get(const char* name, double* ms, long long* launches, double* units);

/* ----------------------------------------------------------------------------- (1) host ABI */

/* reference: stardist3d_lib.h:55-65 (identical signature) */
void _LIB_non_maximum_suppression_sparse(
    const float* scores, const float* dist, const float* points,
    const int n_polys, const int n_rays, const int n_faces,
    const float* verts, const int* faces,
    const float threshold, const int use_bbox, const int use_kdtree,
    const int verbose, bool* result);

/* reference: stardist3d_lib.h:67-82 (identical signature) */
void _LIB_polyhedron_to_label(
    const float* dist, const float* points, const float* verts, const int* faces,
    const int n_polys, const int n_rays, const int n_faces, const int* labels,
    const int nz, const int ny, const int nx,
    const int render_mode, const int verbose,
    const int use_overlap_label, const int overlap_label,
    int* result);

/* 2D analogue of c_non_max_suppression_inds (stardist2d.cpp:390): dist[n_polys*n_rays],
 * points[n_polys*2] (y,x), sorted by descending score; result[n_polys] = kept. Returns status. */
int _LIB_non_maximum_suppression_2d(
    const float* dist, const float* points, const int n_polys, const int n_rays,
    const float threshold, const int use_bbox, const int use_kdtree, const int verbose,
    bool* result);

/* 2D analogue of polygons_to_label_coord (geom2d.py:149-166): coord[n_polys*2*n_rays] as
 * (n_polys, 2, n_rays) float32 with row 0 = y(r), row 1 = x(c); polygons are painted in the
 * given order (later overwrite earlier) with value labels[i]+1. result is int32[ny*nx]. */
int _LIB_polygons_to_label_2d(
    const float* coord, const int* labels, const int n_polys, const int n_rays,
    const int ny, const int nx, int* result);

/* ----------------------------------------------------------------------------- (2) device ABI */

/* 2D NMS, all pointers are device pointers; d_keep is uint8[n_polys]. */
int sdb_nms2d(const float* d_dist, const float* d_points, int n_polys, int n_rays,
              float threshold, int use_bbox, int use_kdtree, int verbose,
              unsigned char* d_keep, sdb_stream_t stream);
/* sdb_nms2d + the ordered list of survivors: d_kept_index[0..*h_n_kept) = indices with keep == 1, ascending
 * (= descending score); d_keep may be NULL.  One 4-byte device->host read-back, stream synchronised on return. */
int sdb_nms2d_survivors(const float* d_dist, const float* d_points, int n_polys, int n_rays,
                        float threshold, int use_bbox, int use_kdtree, int verbose,
                        unsigned char* d_keep, int* d_kept_index, int* h_n_kept, sdb_stream_t stream);
/* Paint order of polygons_to_label (geom2d.py:191-197) for survivors listed by descending score (ties in the
 * order np.argsort(prob, kind='stable')[::-1] leaves them): rank[i] = position of i in
 * np.argsort(prob, kind='stable'), id_by_rank[rank[i]] = i + 1. */
int sdb_paint_order_2d(const float* d_prob_desc, int n, int* d_rank, int* d_id_by_rank, sdb_stream_t stream);
/* Pre-filter of the 2D pair test (csrc/polyfast.cuh): 0 = exact Clipper-equivalent sweep on every pair,
 * 1 (default) = closed-form overlap integral with a conservative bound first, exact sweep for the rest,
 * 2 = verify: both on every pair, disagreements are counted.  Results are identical in all modes.
 * stats: out4 = {pairs tested, pairs that needed the exact sweep, verify mismatches, nms calls}. */
int sdb_nms2d_set_filter(int mode);
/* 1 (default): frontier rounds >= 1 run inside one cooperative kernel with on-device termination; 0: host-driven rounds.
 * Results are identical. */
int sdb_nms2d_set_tail(int on);
void sdb_nms2d_filter_stats(unsigned long long* out4, int reset);

/* paint polygons (geom2d.py:149-197): d_rank[i] = paint rank of polygon i (0 = painted first;
 * a pixel takes the covering polygon of highest rank), d_id_by_rank[r] = value written for
 * rank r (reference: original index + 1); d_out int32[ny*nx] is zeroed first. */
int sdb_polygons_to_label_2d(const float* d_coord, const int* d_rank, const int* d_id_by_rank,
                             int n_polys, int n_rays, int ny, int nx, int* d_out,
                             sdb_stream_t stream);

/* dist_to_coord (geom2d.py:130-146): coord[n,2,R] = f32(dist * sincos(phi_k) in f64) * scale + points.
 * d_sincos = float64[2*R] = [sin(phi_k) | cos(phi_k)] evaluated by the caller with numpy (same
 * libm/SIMD routine as the reference); d_points float64[n,2] (y,x) (integer pixel centres, or
 * already multiplied by the predict_instances(scale=...) rescale, model2d.py:539-546). */
int sdb_dist_to_coord_2d(const float* d_dist, const double* d_points, int n_polys, int n_rays,
                         const double* d_sincos, double scale_y, double scale_x,
                         float* d_coord, sdb_stream_t stream);

/* 3D label painting on device arrays (see _LIB_polyhedron_to_label): polyhedra in painting order
 * (descending probability, geom3d.py:176-180), first cover wins; d_labels[n_polys] must be non-zero. */
int sdb_polyhedron_to_label(const float* d_dist, const float* d_points, const float* d_verts,
                            const int* d_faces, int n_polys, int n_rays, int n_faces,
                            const int* d_labels, int nz, int ny, int nx, int render_mode,
                            int use_overlap_label, int overlap_label, int* d_result,
                            sdb_stream_t stream);

/* sphere culling of bounding-box voxels in the 3-D label rendering (1 = default); label maps are identical either way */
int sdb_label3d_set_cull(int on);

/* k_heavy volume stages (S3 kernel / S4 hull intersection): 0 (default) = per-(k, j) plane scaling as validated on B200,
 * 1 = planes scaled once per pair + early-out for non-cutting planes (sd3::face_cone_volume_n; bit-identical results on the
 * host build, not yet run on a GPU -- experimental until then). */
int sdb_nms3d_set_variant(int norm_planes);
/* 1 (default): a rigorous lower bound of the kernel-intersection volume (ray-wise distances to the 2F planes from the centre
 * midpoint, fan of tetrahedra) decides `iou > threshold` of stage S3 before the volume itself is computed; 0: always the
 * volume.  Decisions are identical (margin 1e-5 relative).  The same switch governs the two-sided fan bounds of S4 / S3 in the
 * second launch and the direction bins of S5 (2: bounds on the coarse ray fan, without the extra ray per face). */
int sdb_nms3d_set_s3_bound(int on);
/* 1 (default): the heavy stages run as three launches per round -- S3 for all pairs, hull facets of the polyhedra that stay open
 * (one warp each), S4 + S5 with those facets; 0: one launch doing everything per pair.  Decisions identical. */
int sdb_nms3d_set_split(int on);

/* relabel_sequential on a device label map (stardist/matching.py:319-406; callers model3d.py:645, base.py:959):
 * labels occurring in d_labels[n] (values in [0, max_label], 16-byte aligned) are renumbered offset, offset+1, ...
 * in ascending order, 0 stays 0, in place.  d_forward_map: scratch of max_label + 3 ints; on return [0..max_label]
 * holds the forward map (matching.py:399-401).  *h_count = number of distinct non-zero labels.  Negative labels
 * are an error like in the reference (matching.py:374).  Synchronises the stream once (count read-back). */
int sdb_relabel_sequential(int* d_labels, long long n, int max_label, int offset, int* d_forward_map,
                           int* h_count, sdb_stream_t stream);

/* 3D NMS on device arrays (see _LIB_non_maximum_suppression_sparse); d_keep is uint8[n_polys]. */
int sdb_nms3d(const float* d_dist, const float* d_points, const float* d_verts, const int* d_faces,
              int n_polys, int n_rays, int n_faces, float threshold, int use_bbox, int use_kdtree,
              int verbose, unsigned char* d_keep, sdb_stream_t stream);

/* threshold + border mask + compaction + sort (nms.py:6-17, base.py:606-610, nms.py:167):
 * prob is [H*W] (2D) / [D*H*W] (3D) float32; candidates are pixels with prob > thresh that lie at
 * least b_lo/b_hi pixels inside along each axis and inside valid_* (un-padded) extents.
 * Output sorted by (prob desc, flat index desc) == np.argsort(prob, kind='stable')[::-1].
 * d_count receives the number of candidates (also returned through *h_count after a sync). */
int sdb_threshold_sort(const float* d_prob, int ndim, const int* shape, const int* valid_shape,
                       const int* b_lo, const int* b_hi, float prob_thresh,
                       int* d_sorted_index, float* d_sorted_prob, int capacity, int* h_count,
                       sdb_stream_t stream);

/* gather: out_dist[r,:] = max(1e-3, dist[idx[r],:]), out_points[r,:] = unravel(idx[r]) * grid */
int sdb_gather_candidates(const float* d_dist, const int* d_index, int n, int n_rays, int ndim,
                          const int* shape, const int* grid, float* d_out_dist,
                          float* d_out_points, sdb_stream_t stream);

/* sparse candidate store (large volumes: the dense dist map is never written, SURVEY H7 / base.py:580-593) */
int sdb_count_above(const float* d_prob, long long n, float thresh, int* h_count, sdb_stream_t stream);
int sdb_store_rows_above(const float* d_prob, const float* d_dist, long long n, int n_rays, float thresh, long long flat0,
                         int row0, int capacity, float* d_store, int* d_slot, sdb_stream_t stream);
int sdb_gather_candidates_slots(const float* d_store, const int* d_slot, const int* d_index, int n, int n_rays, int ndim,
                                const int* shape, const int* grid, float* d_out_dist, float* d_out_points, sdb_stream_t stream);

/* ---- U-Net forward building blocks (NHWC float32; see stardist_b200/csrc/unet*.cu) ---- */

/* 3x3 'same' convolution + bias + optional ReLU.  The input may be the channel concatenation
 * [nearest-upsample2x(in_lo), in_skip] (csbdeep unet_block decoder): pass in_lo=NULL,
 * cin_lo=0 for a plain convolution.  weights are Keras layout (3,3,Cin,Cout), Cin = cin_lo+cin_skip. */
int sdb_conv3x3_2d(const float* d_in, const float* d_in_lo, int n, int h, int w, int cin_skip,
                   int cin_lo, const float* d_weight, const float* d_bias, int cout, int relu,
                   float* d_out, sdb_stream_t stream);

int sdb_maxpool2x2_2d(const float* d_in, int n, int h, int w, int c, float* d_out, sdb_stream_t stream);

/* N-d variants (NDHWC): kz = 1 -> 2-D (d == 1), kz = 3 -> 3x3x3 with Keras kernels (3,3,3,Cin,Cout);
 * (uz,uy,ux) = nearest up-sampling factors of the optional low-resolution source; pooling (pz,py,px). */
int sdb_conv3_nd(const float* d_in, const float* d_in_lo, int n, int d, int h, int w, int cin_skip,
                 int cin_lo, int uz, int uy, int ux, const float* d_weight, const float* d_bias, int cout,
                 int kz, int relu, float* d_out, sdb_stream_t stream);
int sdb_maxpool_nd(const float* d_in, int n, int d, int h, int w, int c, int pz, int py, int px, float* d_out, sdb_stream_t stream);
/* ResNet backbone (model3d.py:402-447): convolution with arbitrary kernel extent and strides under TensorFlow's
 * padding='same' rule (asymmetric: pad_before = pad_total / 2), fp32 CUDA cores; residual add (+ ReLU). */
int sdb_conv_generic_nd(const float* d_in, int n, int d, int h, int w, int cin, const float* d_w, const float* d_b, int cout,
                        int kz, int ky, int kx, int sz, int sy, int sx, int relu, float* d_out, sdb_stream_t stream);
int sdb_add_act(const float* d_a, const float* d_b, long long n, int relu, float* d_out, sdb_stream_t stream);
/* multi-class head (model2d.py:339-347): prob_class = softmax over n_out = n_classes + 1 outputs of a 1x1 convolution of
 * fp32 features [npix, cfeat] with weights [cfeat][n_out]; sdb_merge_split: split fp16 planes (hi, lo) -> fp32. */
int sdb_class_head(const float* d_feat, long long npix, int cfeat, const float* d_w, const float* d_b, int n_out, float* d_out,
                   sdb_stream_t stream);
int sdb_merge_split(const void* d_hi, const void* d_lo, long long n, float* d_out, sdb_stream_t stream);

/* 1x1 heads: prob = sigmoid(x.Wp+bp) [npix], dist = x.Wd+bd [npix*n_rays] */
int sdb_heads_2d(const float* d_feat, long long npix, int cfeat, const float* d_wp, const float* d_bp,
                 const float* d_wd, const float* d_bd, int n_rays, float* d_prob, float* d_dist,
                 sdb_stream_t stream);

/* ---- tcgen05 / TMA / TMEM U-Net path (stardist_b200/csrc/unet_tc.cu) -------------------------
 * Activations are two fp16 planes (hi, lo) per tensor, hi = fp16(v), lo = fp16(v - hi); weights are
 * split the same way into [9][cout][cin] planes by sdb_split_weights.  fp32 accumulation in TMEM. */
int sdb_conv3x3_tc(const void* src0_hi, const void* src0_lo, int c_src0, const void* src1_hi, const void* src1_lo,
                   int c_src1, int n, int h, int w, const void* w_hi, const void* w_lo, float w_scale, const float* d_bias, int cout,
                   int relu, int up2x, void* out_hi, void* out_lo, sdb_stream_t stream);
/* 1x1 heads on the tensor cores (one tap, K = cfeat): head weights [1][np][cfeat] split fp16, row 0 = prob,
 * rows 1..n_rays = dist, zero padded to np in {48, 80, 112, 144}; outputs fp32 */
int sdb_heads_tc(const void* f_hi, const void* f_lo, int cfeat, int n, int h, int w, const void* w_hi, const void* w_lo,
                 float w_scale, const float* d_bias, int np, int n_rays, float* d_prob, float* d_dist, sdb_stream_t stream);
/* Last convolution of the backbone fused with the 1x1 heads (model2d.py:329-337: features conv3x3 + ReLU, then
 * Conv2D(1,1,sigmoid) -> prob and Conv2D(n_rays,1,linear) -> dist): src as in sdb_conv3x3_tc, cout = 128,
 * heads_w fp32 [128][36] (columns 0..n_rays-1 = dist kernels, column 32 = prob kernel, rest 0), heads_b fp32 [36];
 * n_rays <= 32.  Outputs prob [n,h,w] and dist [n,h,w,n_rays] fp32; the feature map is not materialised. */
int sdb_conv3x3_heads_tc(const void* src0_hi, const void* src0_lo, int c_src0, const void* src1_hi, const void* src1_lo,
                         int c_src1, int n, int h, int w, const void* w_hi, const void* w_lo, float w_scale, const float* d_bias,
                         int relu, const float* d_heads_w, const float* d_heads_b, int n_rays, float* d_prob, float* d_dist,
                         sdb_stream_t stream);
/* profiling aid: time persistent CTAs that only TMA-load (rows x 130 px x box_c channel) boxes of an [1,h,w,c] fp16
 * tensor (no MMA), loads_per_tile boxes per 128-pixel tile; ms_out = milliseconds per pass */
int sdb_tma_probe(const void* d_act, int h, int w, int c, int box_c, int rows, int loads_per_tile, int reps, float* ms_out, sdb_stream_t stream);
/* profiling aid: u64 [148][8] wait-cycle counters written by the halo-reuse conv (NULL disables) */
int sdb_tc_set_debug(void* d_buf);
/* 3-D U-Net on the tensor cores (model3d.py:360-399): 3x3x3 convolution of ONE volume [d,h,w,c] in split fp16 planes
 * (k_conv_tc4 with the z planes as the tensor map's image axis; weights [27][cout][cin] from sdb_split_weights_3d,
 * tap = dz*9 + dy*3 + dx; up2x: 0 none, 2 = nearest 2x2x2 up-sampling written by the epilogue), max-pooling on split
 * planes, and the fp32 -> split conversion behind the CUDA-core stem. */
int sdb_conv3x3x3_tc(const void* src0_hi, const void* src0_lo, int c_src0, const void* src1_hi, const void* src1_lo,
                     int c_src1, int d, int h, int w, const void* w_hi, const void* w_lo, float w_scale, const float* d_bias, int cout,
                     int relu, int up2x, void* out_hi, void* out_lo, sdb_stream_t stream);
int sdb_split_weights_3d(const float* d_w, int cin, int cout, float w_scale, void* w_hi, void* w_lo, sdb_stream_t stream);
int sdb_maxpool3d_split(const void* in_hi, const void* in_lo, int d, int h, int w, int c, int pz, int py, int px,
                        void* out_hi, void* out_lo, sdb_stream_t stream);
int sdb_split_f32(const float* d_in, long long n, void* out_hi, void* out_lo, sdb_stream_t stream);
/* kernel variant of sdb_conv3x3_tc: 0 (default) = auto, 1 = one 8x16 tile per CTA, 3 = persistent CTAs with
 * double-buffered TMEM accumulators and merged hi/lo weight tile, 4 = 3 + halo reuse (one box load per
 * 32-channel block, taps as shifted descriptors).  Results are identical up to fp32 summation order. */
int sdb_tc_set_variant(int variant);
/* 1 (default): the small split-fp16 products (lo*Whi, hi*Wlo) accumulate in their own TMEM columns, the main accumulator
 * takes one truncating tensor-core add per k-step; 0: all products into one accumulator (the round-1 scheme, kept for A/B
 * error measurements: tests/tools/tc_split_error.py).  Applies to every tcgen05 convolution except k_conv_tc4<128>. */
int sdb_tc_set_split_acc(int on);
int sdb_tc_error_check(sdb_stream_t stream);
/* w_scale: power of two the weights are multiplied by before the split (undone on the accumulator) */
int sdb_split_weights(const float* d_w, int cin, int cout, float w_scale, void* w_hi, void* w_lo, sdb_stream_t stream);
int sdb_stem_split(const float* d_in, int n, int h, int w, int cin, const float* d_w, const float* d_b, int cout, int relu,
                   void* out_hi, void* out_lo, sdb_stream_t stream);
int sdb_maxpool_split(const void* in_hi, const void* in_lo, int n, int h, int w, int c, void* out_hi, void* out_lo, sdb_stream_t stream);
int sdb_heads_split(const void* f_hi, const void* f_lo, long long npix, int cfeat, const float* d_wp, const float* d_bp,
                    const float* d_wd, const float* d_bd, int n_rays, float* d_prob, float* d_dist, sdb_stream_t stream);

/* ---- predict_instances_big on the device (stardist/big.py:340-413 filter_objects, :319-326 write; base.py:959) ---- */
/* per-label bounding boxes of a C-contiguous int32 label tile (ndim 2 or 3): d_bbox int32[(max_label+1)*6] =
 * {min0,min1,min2,max0,max1,max2} (inclusive; 2-D tiles use columns 1,2,4,5), absent labels keep {INT_MAX.., -1..};
 * d_bad[0] != 0 when a value lies outside [0, max_label].  Replaces skimage.measure.regionprops (big.py:373). */
int sdb_label_bbox(const int* d_labels, int ndim, const int* shape, int max_label, int* d_bbox, int* d_bad, sdb_stream_t stream);
/* in place labels[i] = lut[labels[i]] for non-zero entries (foreign objects -> 0, kept objects -> 1..n_kept) */
int sdb_label_remap(int* d_labels, long long n, const int* d_lut, sdb_stream_t stream);
/* BlockND.write with the running label offset folded in: dst[origin+idx] = tile[idx] + add where tile[idx] > 0 */
int sdb_label_write(const int* d_tile, int ndim, const int* tile_shape, int add, int* d_dst, const int* dst_shape,
                    const int* origin, sdb_stream_t stream);

/* ---- the steps in front of the network: csbdeep.utils.normalize / PercentileNormalizer and the `scale=` zoom ---- */
/* exact order statistics of the region [0,valid) of a C-contiguous float32 array: h_out[q] = sorted(region)[ranks[q]], <= 8 ranks */
int sdb_select_ranks(const float* d_x, int ndim, const int* shape, const int* valid, const long long* ranks, int n_ranks,
                     float* h_out, sdb_stream_t stream);
/* in place x = (x - mi) / den in float32 (den = ma - mi + eps computed by the caller in float32), optional clip to [0,1] */
int sdb_normalize_mi_ma(float* d_x, long long n, float mi, float den, int clip, sdb_stream_t stream);
/* scipy.ndimage.zoom(x, zoom, order=1) (stardist/models/base.py:735) for ndim 2 / 3; out_shape = round(in_shape * zoom) */
int sdb_zoom_linear(const float* d_in, int ndim, const int* in_shape, const int* out_shape, float* d_out, int round_int,
                    double int_lo, double int_hi, sdb_stream_t stream);   /* round_int: integer source image -> scipy's rounding */
/* numpy.pad(mode='reflect') at the end of each spatial axis of a channels-last array (StarDistPadAndCropResizer.before) */
int sdb_pad_reflect_end(const float* d_in, int ndim, const int* in_shape, const int* out_shape, int channels, float* d_out,
                        sdb_stream_t stream);

/* ---- the network boundary as C entry points (keras_model.predict, stardist/models/base.py:408-410) ----
 * A network object is built from the configuration fields that define the graph (model2d.py:310-349 / model3d.py:360-399) and
 * the Keras kernels / biases (HOST pointers, layout (k..., Cin, Cout) / (Cout,)) in the order sdb_unet_layer_name(cfg, i)
 * reports for i in [0, sdb_unet_layer_count(cfg)): the convolution layers in topology order, then "prob", "dist".
 * _LIB_unet_forward_2d/3d take a DEVICE input [h, w, Cin] / [d, h, w, Cin] float32 (normalised, padded to multiples of
 * 2^depth) and write prob [h, w] / [d, h, w] and dist [..., n_rays] (raw head outputs: sigmoid applied to prob, dist
 * not yet clamped to 1e-3).  Default architecture family only (U-Net, 3^d kernels, pool 2, ReLU, no batch norm, grid 1);
 * sdb_unet_create returns NULL (see sdb_last_error) for anything else. */
typedef struct sdb_unet sdb_unet;
typedef struct {
  int ndim, n_channel_in, n_rays, unet_n_depth, unet_n_filter_base, unet_n_conv_per_depth, net_conv_after_unet;
  int grid[3];
} sdb_unet_config;
int sdb_unet_layer_count(const sdb_unet_config* cfg);
const char* sdb_unet_layer_name(const sdb_unet_config* cfg, int i);
sdb_unet* sdb_unet_create(const sdb_unet_config* cfg, const float* const* kernels, const float* const* biases);
void sdb_unet_destroy(sdb_unet* net);
int _LIB_unet_forward_2d(sdb_unet* net, const float* d_x, int h, int w, float* d_prob, float* d_dist, sdb_stream_t stream);
int _LIB_unet_forward_3d(sdb_unet* net, const float* d_x, int d, int h, int w, float* d_prob, float* d_dist, sdb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STARDIST_B200_H */
