/*
 * smesh.h -- C ABI of libsmesh_b200.so: the B200 (sm_100a) implementation of the two semantic-meshes hot paths.
 *
 * This is the drop-in boundary. The reference exposes these paths through two Boost.Python modules
 * (python/semantic_meshes/src/Render.cu, python/semantic_meshes/src/Fusion.cu); every entry point below names the
 * reference interface it replaces (paths relative to the reference checkout; tt/ = extern/template-tensors/include/
 * template_tensors/). INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller owns every buffer (the Python layer holds them as torch tensors); the library keeps no state between
 *     calls except the per-thread error string and the per-kernel launch configuration, so it is safe to use from
 *     several host threads / streams;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and never synchronises the device;
 *   - return value 0 = success; anything else is an smesh_status and smesh_last_error() describes it
 *     (SMESH_ERR_INVALID_ARGUMENT maps to the reference's std::invalid_argument -> Python ValueError);
 *   - image layout is the reference's: (W, H[, C]) row-major, pixel (x, y) at x*H + y
 *     (python/semantic_meshes/include/Renderer.h:29).
 */
#ifndef SMESH_H
#define SMESH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum
{
  SMESH_OK = 0,
  SMESH_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument in the reference (Mesh.h:68-74, Fusion.cu:122-125) */
  SMESH_ERR_CUDA = 2,             /* a CUDA runtime call or kernel launch failed */
  SMESH_ERR_UNSUPPORTED = 3       /* shape outside what the kernels support (see each function) */
} smesh_status;

/* Aggregator kinds; names and arithmetic of python/semantic_meshes/src/Fusion.cu:46-92. */
typedef enum
{
  SMESH_KIND_SUM = 0,
  SMESH_KIND_SUMMAX = 1,
  SMESH_KIND_MUL = 2
} smesh_kind;

/* Element types accepted for primitive-index images (python/semantic_meshes/include/Common.h:5-12). */
typedef enum
{
  SMESH_ID_U32 = 0,
  SMESH_ID_I32 = 1,
  SMESH_ID_U64 = 2,
  SMESH_ID_I64 = 3
} smesh_id_dtype;

/* Last error message of the calling host thread ("" if none). */
const char* smesh_last_error(void);

/* Library / build identification, e.g. "smesh_b200 0.1 sm_100a". */
const char* smesh_version(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Rasterizer: replaces Renderer<T>::render (python/semantic_meshes/include/Renderer.h:25-43) ->
 * TriangleRenderer::render (include/semantic_meshes/render/TriangleRenderer.h:63-89) ->
 * kernel_DeviceMutexRasterizer_nthreadsperprimitive (tt/geometry/render/DeviceMutexRasterizer.h:14-57) +
 * Triangle::{precompute,intersect,rasterize} (tt/geometry/render/primitives/Triangle.h:47-164).
 * ------------------------------------------------------------------------------------------------------------- */

/*
 * Prepared mesh. smesh_raster_mesh_build turns the arrays TriangleRenderer's ctor uploads (TriangleRenderer.h:30-39) into
 * one device blob, once per mesh: faces sorted along a Morton curve of their centroids and cut into units of 32 (one per
 * lane of a warp); a unit is ONE contiguous 1536-byte block holding, per face, its three vertices and its original index
 * (a view fetches it with a single bulk copy, no index -> vertex gathers), with a bounding sphere (a view skips the units
 * that are provably behind the camera or far outside the image); every face is tagged "well shaped" (sine of the smallest
 * angle >= 0.1) or not. 48 bytes per face. The index image still reports ORIGINAL face indices. Results do not depend on
 * the order of the faces.
 *   verts   float32[V][3], faces int32[F][3] with 0 <= index < V   (F < 2^32 - 1, V < 2^31)
 *   mesh_out / temp   256-byte aligned device buffers of the sizes smesh_raster_mesh_bytes reports; temp is scratch for
 *           the build only (sort keys), mesh_out is what smesh_raster_render takes
 */
int smesh_raster_mesh_bytes(int64_t V, int64_t F, size_t* mesh_bytes_host, size_t* temp_bytes_host);
int smesh_raster_mesh_build(const float* verts, int64_t V, const int32_t* faces, int64_t F, void* mesh_out, size_t mesh_bytes,
                            void* temp, size_t temp_bytes, void* stream);

/* Bytes of device scratch smesh_raster_render needs for a mesh of V vertices / F triangles at resolution W x H
 * (per-view ray tables, packed 64-bit depth|index buffer, cluster and large-triangle queues). */
int smesh_raster_workspace_bytes(int64_t V, int64_t F, int W, int H, size_t* bytes_host);

/*
 * Render one view: per pixel the index of the nearest hit triangle and its camera-space depth z.
 *   mesh    the blob smesh_raster_mesh_build filled for (V, F)
 *   R_host  float[9] row-major rotation, t_host float[3]   (Camera::extr, include/semantic_meshes/render/Camera.h:12)
 *   f_host  double[2] focal lengths, c_host double[2] principal point (Camera::intr; the Python Camera rounds its
 *           inputs to float first and then widens, python/semantic_meshes/include/Camera.h:19-54 - the caller does it)
 *   W, H    Camera::resolution (1 <= W, H <= 65536)
 *   workspace  device scratch of smesh_raster_workspace_bytes(V, F, W, H) bytes (ray tables, packed depth buffer, unit
 *           list, big-triangle queue): every view initialises what it uses, nothing is carried from view to view; keep
 *           it for the next views of the same mesh and resolution; one workspace serves one stream at a time
 *   idx_out uint32[W][H] (0xFFFFFFFF where nothing is hit), depth_out float32[W][H] (+inf where nothing is hit; may be
 *           NULL when the caller only fuses the view: the depth image is then not written)
 * Results are bit-identical to the reference kernel as compiled by nvcc 12.9 for sm_100a, with the one documented
 * strengthening that exact depth ties go to the lowest triangle index (the reference is order-dependent there).
 */
int smesh_raster_render(const void* mesh, size_t mesh_bytes, int64_t V, int64_t F, const float* R_host, const float* t_host,
                        const double* f_host, const double* c_host, int W, int H, void* workspace, size_t workspace_bytes,
                        uint32_t* idx_out, float* depth_out, void* stream);

/*
 * smesh_raster_render that also leaves the per-face pixel counts of the view (Mesh.h:90-93) in `counts` (uint32[F], the
 * tagged counters described under "Label fusion" below) with epoch count_epoch in 1..255: the first half of
 * ModelAggregator::add for an aggregator over the faces of this mesh, fused into the pass that writes the index image.
 * Follow with smesh_fuse_scatter(ids32 = idx_out, same counts, same count_epoch).
 */
int smesh_raster_render_counted(const void* mesh, size_t mesh_bytes, int64_t V, int64_t F, const float* R_host,
                                const float* t_host, const double* f_host, const double* c_host, int W, int H, void* workspace,
                                size_t workspace_bytes, uint32_t* idx_out, float* depth_out, uint32_t* counts,
                                uint32_t count_epoch, void* stream);

/*
 * Texel renderer: replaces semantic_meshes::render::TexturedTriangleRenderer
 * (include/semantic_meshes/render/TexturedTriangleRenderer.h), the `render.texels(...)` of
 * python/semantic_meshes/src/Render.cu:20-23: primitives are the TEXELS of per-triangle textures instead of the triangles.
 *
 * smesh_texels_prepare = its constructor (:86-176), on the HOST like the reference's (its decisions hang on host float
 * arithmetic, glibc acosf included): per triangle the texture resolution from the largest projected area over all cameras
 * (tri_res_host uint32[F]), the corner that becomes the texture origin (faces_host int32[F][3] is REORDERED IN PLACE) and
 * the first texel of each triangle (first_texel_host uint32[F]); *n_texels_host = getPrimitivesNum().
 *   cameras: R_host float[n][9] row-major, t_host float[n][3], f_host / c_host double[n][2], resolution_host int32[n][2]
 * Build the prepared mesh (smesh_raster_mesh_build) from the REORDERED faces, upload tri_res / first_texel, then
 * smesh_raster_render_texels renders a view like smesh_raster_render, idx_out holding texel indices
 * (TexturedTriangle::getTexelIndex, :32-41).
 */
int smesh_texels_prepare(const float* verts_host, int64_t V, int32_t* faces_host, int64_t F, int n_cameras, const float* R_host,
                         const float* t_host, const double* f_host, const double* c_host, const int32_t* resolution_host,
                         float texels_per_pixel, uint32_t* tri_res_host, uint32_t* first_texel_host, uint64_t* n_texels_host);
int smesh_raster_render_texels(const void* mesh, size_t mesh_bytes, int64_t V, int64_t F, const uint32_t* tri_res,
                               const uint32_t* first_texel, const float* R_host, const float* t_host, const double* f_host,
                               const double* c_host, int W, int H, void* workspace, size_t workspace_bytes, uint32_t* idx_out,
                               float* depth_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Label fusion: replaces ModelAggregator::{add1,add2,get,reset} (python/semantic_meshes/include/Fusion.h:42-76) ->
 * semantic_meshes::ModelAggregator::{add,get,reset} (include/semantic_meshes/fusion/Mesh.h:57-133) with the chains
 * of python/semantic_meshes/src/Fusion.cu:46-92.
 *
 * Accumulator: float32[P][Cpad] with Cpad = smesh_fuse_padded_classes(C) (rows padded to 16 bytes so a row is
 * updated with 128-bit reductions); padding columns stay 0. For SMESH_KIND_MUL it holds -log p (tt/numeric/LogProb.h).
 * reset() of the reference (Mesh.h:119-122) = zero-fill of this buffer by the caller.
 * ------------------------------------------------------------------------------------------------------------- */

int smesh_fuse_padded_classes(int C);

/*
 * Per-view pixel counters. counts is uint32[P] device scratch owned by the caller; every call that counts a view takes
 * a `count_epoch`:
 *   1..255  tagged mode (normal): a word holds (epoch << 24) | n. The view first raises the words it touches to its own
 *           epoch and then counts, so nothing has to be cleared between views; the caller hands out strictly increasing
 *           epochs and zero-fills counts before starting over at 1 (and before the first use). Needs n_pix < 2^24.
 *   0       untagged mode: counts must be all-zero on entry and is all-zero again on completion (an extra clear launch);
 *           for images of 2^24 pixels or more.
 *
 * smesh_fuse_add: one view into the accumulator = ModelAggregator::add (Mesh.h:65-107). Launches: per-face pixel count
 * (Mesh.h:90-93) and the gated, weighted scatter (Mesh.h:94-106).
 *
 * Pixels are addressed by a flat index i = outer*n_inner + inner (n_pix = n_outer*n_inner); the caller picks
 * outer/inner so that the probability image is contiguous in that order:
 *   probs    float32[n_pix][C]  (class stride 1; 16-byte aligned for the fast path)
 *   ids      element (outer, inner) at ids[outer*ids_stride_outer + inner*ids_stride_inner] (strides in elements),
 *            type id_dtype; a pixel is background unless 0 <= id < P (Mesh.h:95)
 *   weights  NULL (= 1.0f everywhere, Mesh.h:109-117) or float32 contiguous in the same flat order
 *            (w_stride_inner == 1, w_stride_outer == n_inner)
 *   ids32    uint32[n_pix] scratch (flat-order copy of the ids, 0xFFFFFFFF for background; untouched when the ids are
 *            already 32-bit and flat)
 *   acc      float32[P][Cpad]
 *   iew      images_equal_weight (Mesh.h:57,102)
 * SMESH_ERR_UNSUPPORTED if C > 4096 or P >= 2^32 - 1.
 */
int smesh_fuse_add(int kind, const void* ids, int id_dtype, int64_t ids_stride_outer, int64_t ids_stride_inner,
                   const float* probs, const float* weights, int64_t w_stride_outer, int64_t w_stride_inner,
                   int64_t n_outer, int64_t n_inner, int C, int64_t P, float iew, uint32_t* counts, uint32_t count_epoch,
                   uint32_t* ids32, float* acc, void* stream);

/*
 * The stages of smesh_fuse_add as separate calls (a caller that adds the same index image with several predictions can
 * count once; bench.py times the scatter stage alone with them). Pixels in flat order:
 *   smesh_fuse_count    per-face pixel count (Mesh.h:90-93) of ids (any id_dtype, strided) into counts[P]; ids32_out
 *                       (may be NULL) receives the flat uint32 copy, background and out-of-range ids as 0xFFFFFFFF
 *   smesh_fuse_scatter  gate + weight + accumulate (Mesh.h:94-106) from flat uint32 ids (anything >= P is background),
 *                       flat probs [n_pix][C], flat weights (NULL = 1) and the counts of the same view / same epoch
 *   smesh_fuse_clear    counts[id] = 0 for every id < P in ids32 (what count_epoch 0 needs after the scatter)
 */
int smesh_fuse_count(const void* ids, int id_dtype, int64_t ids_stride_outer, int64_t ids_stride_inner, int64_t n_outer,
                     int64_t n_inner, int64_t P, uint32_t* counts, uint32_t count_epoch, uint32_t* ids32_out, void* stream);
int smesh_fuse_scatter(int kind, const uint32_t* ids32, const float* probs, const float* weights, int64_t n_pix, int C,
                       int64_t P, float iew, const uint32_t* counts, uint32_t count_epoch, float* acc, void* stream);
int smesh_fuse_clear(const uint32_t* ids32, int64_t n_pix, int64_t P, uint32_t* counts, void* stream);

/*
 * One view's scatter stage with the NEXT view's count stage riding in the same launch (one extra warp per CTA of the ring
 * kernels counts the next index image while the others scatter this view; where the kernel has no spare warp the next
 * count is a launch of its own): what smesh_fuse_add_batch does between its views, for callers that hold the next index
 * image early - a pipeline that renders one view ahead. Flat uint32 ids, tagged epochs only:
 *   counts / count_epoch            this view's counter array and epoch; counted != 0: the counts are already in place
 *                                   (a previous call carried them, or smesh_raster_render_counted), else they are taken first
 *   next_ids32 / next_n_pix         the next view's flat index image (may have another size)
 *   next_counts / next_epoch        the OTHER counter array and an epoch of the other parity
 * On completion next_counts holds the next view's counts: follow with this function again (counted = 1) or with
 * smesh_fuse_scatter(ids32 = next_ids32, counts = next_counts, count_epoch = next_epoch).
 */
int smesh_fuse_scatter_count_next(int kind, const uint32_t* ids32, const float* probs, const float* weights, int64_t n_pix,
                                  int C, int64_t P, float iew, uint32_t* counts, uint32_t count_epoch, int counted,
                                  const uint32_t* next_ids32, int64_t next_n_pix, uint32_t* next_counts, uint32_t next_epoch,
                                  float* acc, void* stream);

/*
 * A batch of B views with identical shapes, view b at ids + b*ids_stride_view (elements), probs + b*probs_stride_view
 * (floats), weights + b*w_stride_view (floats, if weights != NULL). Equivalent to B calls of smesh_fuse_add in order with
 * epochs count_epoch0, count_epoch0 + 1, ... (all <= 255), or all 0.
 * (the ORDER of the float additions into a row is not defined, exactly as between two pixels of one view).
 *   counts2  uint32[2][P]: TWO counter arrays; the view with epoch e counts into array e & 1 (epoch 0: array 0 only).
 *            With tagged epochs and 32-bit flat ids the views are dealt to two lanes - `stream` and a side stream owned by
 *            the library (per host thread and device), forked from and joined to `stream` by events, so the call is
 *            stream-ordered on `stream` and can be captured into a CUDA graph - and each lane runs count + scatter of its
 *            views on its own counter array: one lane's count stage, launch gaps and tails are covered by the other's
 *            scatter. If the side stream does not exist yet while `stream` is being captured (nothing is created during a
 *            capture): one stream, the count stage of view b+1 riding in the scatter launch of view b.
 */
int smesh_fuse_add_batch(int kind, int64_t B, const void* ids, int id_dtype, int64_t ids_stride_view,
                         int64_t ids_stride_outer, int64_t ids_stride_inner, const float* probs,
                         int64_t probs_stride_view, const float* weights, int64_t w_stride_view, int64_t w_stride_outer,
                         int64_t w_stride_inner, int64_t n_outer, int64_t n_inner, int C, int64_t P, float iew,
                         uint32_t* counts2, uint32_t count_epoch0, uint32_t* ids32, float* acc, void* stream);

/*
 * ModelAggregator::get (Fusion.h:72-76): out float32[P][C] = per-face class distribution: the accumulator row
 * (mul: exp(-(acc - min acc)), Fusion.h:96-104), L1-normalised, NaN/Inf -> 0 (Fusion.cu:66-92, Fusion.h:79-95).
 */
int smesh_fuse_get(int kind, const float* acc, int64_t P, int C, float* out, void* stream);

/*
 * What follows get() in the reference's own pipeline.
 *
 * smesh_fuse_labels: per-face label from the distribution smesh_fuse_get returned (dist float32[P][C]), exactly
 * python/scripts/colorize_mesh.py:82-88: -1 ("no annotation") if the row sums to less than dont_care_threshold (0.9 in the
 * script), else the first class of maximal probability (tf.argmax). labels_out int32[P].
 *
 * smesh_fuse_render: ModelRenderer::render (include/semantic_meshes/fusion/Mesh.h:24-43), the gather-back of per-face
 * annotations into an image: out[i] = annotations[ids32[i]] if ids32[i] < P else background, elements of elem_bytes
 * bytes (labels: 4, colours: 3, distributions: 4*C); annotations [P], background one element, out [n_pix], all on the
 * device.
 */
int smesh_fuse_labels(const float* dist, int64_t P, int C, float dont_care_threshold, int32_t* labels_out, void* stream);
int smesh_fuse_render(const void* annotations, int64_t P, int elem_bytes, const uint32_t* ids32, int64_t n_pix,
                      const void* background, void* out, void* stream);

/*
 * Self-test (used by tests/): the rasterizer normalises a pixel's ray with 1 / sqrt(l2), both operations IEEE
 * round-to-nearest (MiscOps.h:125-128), l2 >= 1. The kernels evaluate that with the fast paths of nvcc's own expansion
 * of sqrt.rn and rcp.rn written out without their range checks. This entry compares the two for the n floats whose bit
 * patterns start at first_bits (first_bits + n <= 0x7F800000) and leaves the number of differing results in
 * *mismatches_dev (device uint64): 0 for every float in [1, 2^100).
 */
int smesh_selftest_inv_sqrt(uint32_t first_bits, uint64_t n, uint64_t* mismatches_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * The per-view loop of the reference's scripts, `idx, depth = renderer.render(cam); aggregator.add(idx, probs)`
 * (python/scripts/colorize_mesh.py:60-72), for B views of one resolution enqueued by ONE call: equivalent to
 *   for b in 0 .. B-1:  smesh_raster_render(view b -> idx, depth);  smesh_fuse_add(idx, probs_b, weights_b, epoch0 + b)
 * with the renders up to `ring` views ahead of the fusion on a side stream owned by the library (per host thread and device; forked
 * from and joined to `stream` by events, so the call is stream-ordered on `stream` and can be captured into a CUDA graph;
 * if that stream does not exist yet while `stream` is being captured, everything runs on `stream` in order).
 *   R_host float[B][9], t_host float[B][3], f_host double[B][2], c_host double[B][2]: the cameras (smesh_raster_render)
 *   ring      1 .. SMESH_PIPELINE_MAX_RING index images in flight. With 2 the render of view b+2 has to wait for the fusion
 *             of view b and the two streams fall into lock step (measured 11.0 k against 12.6 k views/s at 2 M triangles,
 *             2048x1024); 4 decouples them
 *   idx_ring  uint32[ring][W*H] device scratch (on return it holds the index images of the last views)
 *   depth_ring float32[ring][W*H] or NULL (no depth images are written)
 *   probs  host array of B device pointers: view b's float32 (W, H, C) contiguous predictions (any order, repeats allowed)
 *   weights NULL, or a host array of B device pointers to float32 (W, H) contiguous weight images
 *   counts2 uint32[2][P] as for smesh_fuse_add_batch, epochs count_epoch0 .. count_epoch0 + B - 1 inside 1 .. 255
 *   (so W*H < 2^24)
 * ------------------------------------------------------------------------------------------------------------- */
#define SMESH_PIPELINE_MAX_RING 8
int smesh_pipeline_views(const void* mesh, size_t mesh_bytes, int64_t V, int64_t F, int64_t B, const float* R_host,
                         const float* t_host, const double* f_host, const double* c_host, int W, int H, void* workspace,
                         size_t workspace_bytes, int ring, uint32_t* idx_ring, float* depth_ring, int kind,
                         const float* const* probs,
                         const float* const* weights, int C, int64_t P, float iew, uint32_t* counts2, uint32_t count_epoch0,
                         float* acc, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* SMESH_H */
