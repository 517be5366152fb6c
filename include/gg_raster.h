/* gg_raster.h -- C ABI of the B200-native differentiable Gaussian-splat rasterizer.
 *
 * This is the drop-in boundary for the hot path of eth-ait/Gaussian-Garments.  In the
 * reference that boundary is the pybind11 module `_C` of the third-party package
 * `diff_gaussian_rasterization_depth_alpha` (imported at
 * /root/reference/gaussian_renderer/__init__.py:16, installed by /root/reference/setup.sh:26-29),
 * which the Python classes `GaussianRasterizationSettings` / `GaussianRasterizer` bind to
 * (/root/reference/gaussian_renderer/__init__.py:39-54 and :103-111).  Each entry point below
 * names the `_C` function (upstream recall) or reference call site it replaces.
 *
 * Conventions
 *   - plain C: device pointers are `void*`/typed pointers to CUDA device memory, sizes are
 *     integers, the stream is a `cudaStream_t` passed as `void*`.  No torch types.
 *   - the library never allocates or frees device memory: every output and workspace is
 *     caller-allocated (the Python host lets torch's caching allocator own them; SURVEY.md 8b
 *     "Ownership").  Workspace sizes come from the gg_*_workspace_bytes() queries.
 *   - every function returns 0 on success, a positive cudaError_t value or a negative GG_E_*
 *     code on failure; gg_last_error() returns a thread-local description.  Nothing throws or
 *     exits across the boundary.
 *   - re-entrant.  Process-wide state: diagnostics (launch counter, optional per-kernel timing events, off the compute
 *     path), environment switches read once, and the forward's fork sets -- one {side stream, two events} per
 *     (device, caller stream), mutex-protected (see gg_forward_color).  `device` is re-asserted with cudaSetDevice
 *     because autograd calls backward on a worker thread (SURVEY.md 8b "Threading / streams").
 *   - all floating tensors are fp32, contiguous, 16-byte aligned base pointers.
 */
#ifndef GG_RASTER_H_
#define GG_RASTER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_ABI_VERSION 1
#define GG_TILE 16

#define GG_E_BADARG (-1)   /* null pointer / inconsistent sizes */
#define GG_E_ALIGN (-2)    /* pointer not 16-byte aligned */
#define GG_E_OVERFLOW (-3) /* num_rendered exceeds the instance workspace capacity given */

/* The scalar part of GaussianRasterizationSettings (gaussian_renderer/__init__.py:39-52).
 * The tensor-valued settings (bg, viewmatrix, projmatrix, campos) stay on the device and are
 * passed by pointer, exactly as the reference passes CUDA tensors. */
typedef struct gg_view {
    int32_t num_gaussians; /* N  = means3D.shape[0]                                   */
    int32_t sh_coeffs;     /* M  = shs.shape[1]; 0 when colors_precomp is used        */
    int32_t sh_degree;     /* D  = raster_settings.sh_degree (active degree, 0..3)    */
    int32_t image_width;   /* raster_settings.image_width                             */
    int32_t image_height;  /* raster_settings.image_height                            */
    float tanfovx;         /* raster_settings.tanfovx                                 */
    float tanfovy;         /* raster_settings.tanfovy                                 */
    float scale_modifier;  /* raster_settings.scale_modifier                          */
    int32_t prefiltered;   /* raster_settings.prefiltered (accepted, unused as upstream) */
    int32_t debug;         /* raster_settings.debug: sync + check after every launch  */
} gg_view;

/* Device-resident inputs of one rasterizer call (the 8 call kwargs of
 * gaussian_renderer/__init__.py:103-111 plus the 4 tensor settings). Unused alternatives NULL. */
typedef struct gg_inputs {
    const float* means3D;        /* [N,3]                                            */
    const float* shs;            /* [N,M,3] or NULL                                  */
    const float* colors_precomp; /* [N,3]   or NULL                                  */
    const float* opacities;      /* [N,1]                                            */
    const float* scales;         /* [N,3]   or NULL                                  */
    const float* rotations;      /* [N,4] wxyz or NULL                               */
    const float* cov3D_precomp;  /* [N,6]   or NULL                                  */
    const float* bg;             /* [3]                                              */
    const float* viewmatrix;     /* [4,4] world_view_transform (scene/cameras.py:59) */
    const float* projmatrix;     /* [4,4] full_proj_transform  (scene/cameras.py:61) */
    const float* campos;         /* [3]   camera_center        (scene/cameras.py:62) */
} gg_inputs;

/* ---- workspace size queries (bytes) ------------------------------------------------------ */
/* geom: per-Gaussian projected records, transient within forward.
 * tile: per-tile counters / offsets; offsets are needed again by backward.
 * image: per-pixel n_contrib + final transmittance, needed again by backward.           */
int gg_forward_workspace_bytes(const gg_view* view, size_t* geom_bytes, size_t* tile_bytes,
                               size_t* image_bytes);
/* keys: transient (tile-bucketed depth keys); records: packed sorted per-instance records,
 * needed again by backward.  `num_rendered` = K.                                          */
int gg_instance_workspace_bytes(int64_t num_rendered, size_t* key_bytes, size_t* record_bytes);
/* per-Gaussian gradient accumulators, transient within backward (gg_backward zero-fills them). */
int gg_backward_workspace_bytes(const gg_view* view, size_t* accum_bytes);

/* ---- forward, stage 1a: replaces the first half of `_C.rasterize_gaussians`
 * (geometry part of preprocessCUDA + InclusiveSum + the num_rendered read-back).
 * Launches: project (3D->2D covariance, cull, tile rectangle, per-tile counts) and the tile
 * scan.  Writes radii[N] (int32, an output tensor of the call).  `num_rendered_host` may be
 * NULL or TWO pinned host words that receive {K, largest per-tile instance count} through an
 * async copy enqueued on `stream` right after the scan: record an event after this call and
 * wait on it before reading.                                                              */
int gg_forward_project(const gg_view* view, const gg_inputs* in, void* geom_ws, void* tile_ws,
                       int32_t* radii, uint32_t* num_rendered_host, int device, void* stream);

/* ---- forward, stage 1b: the colour part of preprocessCUDA (SH deg<=3 -> RGB, +0.5, clamp;
 * or a copy of colors_precomp).  Enqueued behind the K read-back so that the host's wait
 * for K overlaps this kernel (the 57.6 MB SH read at 300k Gaussians).
 * When the preceding gg_forward_project ran on the same (device, stream) the kernel is launched on a library-owned
 * side stream, ordered behind the projection only, so that it runs CONCURRENTLY with the tile scan and the instance
 * emission; the next gg_forward_render on `stream` joins it before the first reader of the colours.  The caller sees
 * plain stream semantics on `stream` (capturable in a CUDA graph).  GG_FWD_FORK=0, debug views and per-kernel timing
 * keep everything on `stream`.                                                                                    */
int gg_forward_color(const gg_view* view, const gg_inputs* in, void* geom_ws, const int32_t* radii,
                     int device, void* stream);

/* ---- forward, stage 2: replaces the second half of `_C.rasterize_gaussians`
 * (duplicateWithKeys + RadixSort + identifyTileRanges + renderCUDA).
 * Launches: instance emit, per-tile depth sort + record packing, front-to-back blend.
 * `instance_capacity` is the K the key/record workspaces were sized for;
 * `max_tile_instances` (second word from gg_forward_project; 0 = unknown) sizes the per-tile
 * sort's shared memory.  Outputs: out_color[3,H,W], out_depth[1,H,W], out_alpha[1,H,W].   */
int gg_forward_render(const gg_view* view, const gg_inputs* in, const void* geom_ws, void* tile_ws,
                      void* key_ws, void* record_ws, int64_t instance_capacity,
                      int64_t max_tile_instances, void* image_ws, const int32_t* radii,
                      float* out_color, float* out_depth, float* out_alpha, int device, void* stream);

/* Same as gg_forward_render, but the colour stage (gg_forward_color must NOT have been called) runs AFTER instance
 * emission and the per-tile sort, behind `color_gate_event` (a cudaEvent_t, may be NULL): the only consumer of the SH
 * coefficients then waits for the previous step's SH-gradient exchange (multi-GPU, dist.py) while projection, emission
 * and sorting of this view are already executing; colours are then scattered into the packed records.
 * color_gate_words (may be NULL) is the graph-replay form of the same gate: three device words {G, X, timeout}; the
 * colour stage first runs a one-warp kernel that takes ticket k = ++G and waits until X >= k - 1, where X is advanced by
 * gg_gate_signal() enqueued behind each SH-gradient exchange -- a dependency that survives CUDA-graph boundaries.   */
int gg_forward_render_late_color(const gg_view* view, const gg_inputs* in, void* geom_ws, void* tile_ws, void* key_ws,
                                 void* record_ws, int64_t instance_capacity, int64_t max_tile_instances, void* image_ws,
                                 const int32_t* radii, float* out_color, float* out_depth, float* out_alpha,
                                 void* color_gate_event, uint32_t* color_gate_words, int device, void* stream);
int gg_gate_signal(uint32_t* gate_words, int device, void* stream);

/* ---- sync-free operation (CUDA graphs): upstream blocks on a D2H copy of num_rendered in the middle of every
 * forward; when the instance workspaces are sized from an earlier call instead, this records on the
 * device whether that was enough: flag2[0] |= 1 if K > instance_capacity (results of that forward are then
 * incomplete), flag2[1] = max(flag2[1], K).  flag2: two device words the caller keeps across calls.              */
int gg_forward_overflow_check(const gg_view* view, const void* tile_ws, int64_t instance_capacity, uint32_t* flag2,
                              int device, void* stream);

/* ---- backward: replaces `_C.rasterize_gaussians_backward`
 * (renderCUDA bwd + computeCov2DCUDA + preprocessCUDA bwd).
 * Upstream gradients dL_dcolor[3,H,W], dL_ddepth[1,H,W], dL_dalpha[1,H,W] (any may be NULL =
 * zeros).  Output gradients (any may be NULL = not wanted): dL_dmeans3D[N,3],
 * dL_dmeans2D[N,3] (side channel, z = 0; consumer scene/gaussian_model.py:410-412),
 * dL_dshs[N,M,3], dL_dcolors_precomp[N,3], dL_dopacities[N,1], dL_dscales[N,3],
 * dL_drotations[N,4], dL_dcov3D[N,6].  accum_ws is scratch (zero-filled here);
 * `instance_capacity` must be the value given to gg_forward_render for `record_ws`.      */
int gg_backward(const gg_view* view, const gg_inputs* in, const void* tile_ws, const void* record_ws,
                int64_t instance_capacity, const void* image_ws, const int32_t* radii, void* accum_ws,
                const float* dL_dcolor,
                const float* dL_ddepth, const float* dL_dalpha, float* dL_dmeans3D, float* dL_dmeans2D,
                float* dL_dshs, float* dL_dcolors_precomp, float* dL_dopacities, float* dL_dscales,
                float* dL_drotations, float* dL_dcov3D, int device, void* stream);

/* ---- replaces `_C.mark_visible` (GaussianRasterizer.markVisible; not called by the
 * reference, kept for API completeness): visible[i] = (view-space z > 0.2).               */
int gg_mark_visible(int32_t num_gaussians, const float* means3D, const float* viewmatrix,
                    const float* projmatrix, uint8_t* visible, int device, void* stream);

/* ---- fused mesh-binding transform ("next" row N1) ------------------------------------------
 * Replaces the torch op chain of /root/reference/scene/mesh_gaussian_model.py:90-128
 * (update_face_coor, get_xyz, get_scaling, get_rotation) and utils/graphics_utils.py:118-137
 * (compute_face_orientation): per-face frames from verts[V,3] / faces[F,3] (int32), then per
 * Gaussian (binding[N] int32 = its face):  xyz = R_f local_xyz * s_f + c_f,
 * scaling = exp(local_log_scaling) * s_f,  rotation = normalize(q_f (x) normalize(local_rotation)) (wxyz).
 * frame_ws (gg_mesh_bind_workspace_bytes) is written by forward and read by backward;
 * frame_grad_ws has the same size.  Backward outputs may be NULL (not wanted); dL_dverts is
 * zero-filled and accumulated here (this is the gradient stage 2 all-reduces).              */
int gg_mesh_bind_workspace_bytes(int32_t num_faces, size_t* frame_bytes);
int gg_mesh_bind_forward(int32_t num_vertices, int32_t num_faces, int32_t num_gaussians, const float* verts,
                         const int32_t* faces, const int32_t* binding, const float* local_xyz,
                         const float* local_log_scaling, const float* local_rotation, void* frame_ws,
                         float* out_xyz, float* out_scaling, float* out_rotation, int device, void* stream);
int gg_mesh_bind_backward(int32_t num_vertices, int32_t num_faces, int32_t num_gaussians, const float* verts,
                          const int32_t* faces, const int32_t* binding, const float* local_xyz,
                          const float* local_log_scaling, const float* local_rotation, const void* frame_ws,
                          void* frame_grad_ws, const float* dL_dxyz, const float* dL_dscaling,
                          const float* dL_drotation, float* dL_dverts, float* dL_dlocal_xyz,
                          float* dL_dlocal_log_scaling, float* dL_dlocal_rotation, int device, void* stream);

/* `_ex` variants: the two remaining forms of the reference's binding.
 *   barycentric[N,3] != NULL  -> AvatarGaussianModel: the local frame is anchored at a*v0 + b*v1 + c*v2 of the bound
 *                                face instead of its centre (get_xyz / get_final_xyz / get_barycentric_3d,
 *                                /root/reference/scene/avatar_gaussian_model.py:140-159; pass `local_xyz` = _xyz or the
 *                                per-frame `local_xyz` of scene/avatar_net.py:82); its gradient goes to the 3 vertices.
 *   face_scaling_remembered[F] != NULL -> get_scaling multiplies with this frozen per-face scale (remember_scaling,
 *                                /root/reference/scene/mesh_gaussian_model.py:98-110) and sends no gradient to the mesh.
 * With both NULL they are the functions above.                                                                      */
int gg_mesh_bind_forward_ex(int32_t num_vertices, int32_t num_faces, int32_t num_gaussians, const float* verts,
                            const int32_t* faces, const int32_t* binding, const float* local_xyz,
                            const float* local_log_scaling, const float* local_rotation, const float* barycentric,
                            const float* face_scaling_remembered, void* frame_ws, float* out_xyz, float* out_scaling,
                            float* out_rotation, int device, void* stream);
int gg_mesh_bind_backward_ex(int32_t num_vertices, int32_t num_faces, int32_t num_gaussians, const float* verts,
                             const int32_t* faces, const int32_t* binding, const float* local_xyz,
                             const float* local_log_scaling, const float* local_rotation, const float* barycentric,
                             const float* face_scaling_remembered, const void* frame_ws, void* frame_grad_ws,
                             const float* dL_dxyz, const float* dL_dscaling, const float* dL_drotation, float* dL_dverts,
                             float* dL_dlocal_xyz, float* dL_dlocal_log_scaling, float* dL_dlocal_rotation, int device,
                             void* stream);

/* ---- on-device visibility ray cast ("next" row N3) ----------------------------------------------
 * Replaces the GPU -> CPU -> Embree -> GPU round trip of get_visible_mask
 * (/root/reference/scene/avatar_gaussian_model.py:227-263, /root/reference/inference.py:285-316): one ray per target
 * from `origin` (the camera centre, device float[3]) along (target - origin)/|target - origin| against the triangle
 * mesh verts[V,3] / faces[F,3] (int32).  primitive_ids[N] (int32) receives what open3d's
 * RaycastingScene.cast_rays()['primitive_ids'] holds -- the index of the closest triangle hit -- with -1 instead of
 * INVALID_ID; t_hit[N] (may be NULL) the hit distance (inf when nothing is hit).  `look_at` (device float[3], e.g. the
 * mesh centroid) only orients the projection grid that accelerates the cast; the answer does not depend on it.
 * The caller forms the masks: primitive_ids == binding (avatar model), geometry_of(primitive_ids) == garment | no hit
 * (inference).  ws / list_capacity from gg_cast_rays_workspace_bytes; force_bruteforce != 0 skips the grid.       */
int gg_cast_rays_workspace_bytes(int32_t num_vertices, int32_t num_faces, size_t* ws_bytes, int64_t* list_capacity);
int gg_cast_rays_from_point(int32_t num_vertices, int32_t num_faces, int32_t num_rays, const float* verts,
                            const int32_t* faces, const float* targets, const float* origin, const float* look_at,
                            void* ws, int64_t list_capacity, int32_t force_bruteforce, int32_t* primitive_ids,
                            float* t_hit, int device, void* stream);

/* ---- in-switch (NVLS) average of the gradient bucket (multi-GPU row 8e) --------------------------------
 * The reference has no distributed code; BASELINE.json's multi-GPU configs shard views one per GPU and exchange only
 * the per-step sum of the parameter gradients.  `multicast_base` is the multicast (multimem) address of a symmetric
 * fp32 buffer every rank has written its local gradients into (the flat bucket the backward kernel's outputs alias);
 * on return -- stream-ordered -- elements [elem_offset, elem_offset + elem_count) of EVERY rank's buffer hold
 * scale * (sum over ranks).  One kernel per rank: multimem.ld_reduce (the switch adds) + multimem.st (the switch
 * broadcasts); rank r reduces the r-th 1/world slice.  signal_pads_dev: device array of world_size pointers to the
 * ranks' uint32 signal pads (zero-initialised; num_blocks * world_size words from pad_slot0 are used -- give concurrent
 * calls on different streams disjoint ranges).  All ranks must call with identical arguments, in the same order.   */
int gg_nvls_allreduce_f32(void* multicast_base, const void* signal_pads_dev, int32_t rank, int32_t world_size,
                          int64_t elem_offset, int64_t elem_count, float scale, int32_t pad_slot0, int32_t num_blocks,
                          int device, void* stream);

/* ---- fused photometric loss ("next" row N2) -------------------------------------------------
 * Replaces l1_loss(image, gt, mask) and ssim(image, gt, mask) of /root/reference/utils/loss_utils.py:17-69
 * as used at s2_registration.py:259-260 / s3_appearance.py:132-133.  image, gt: [3,H,W]; mask: [1,H,W]
 * or NULL.  Forward writes into map_ws: 2 x 64 double accumulator slots at offset 0 (slots 0..63 add up to
 * sum |(image-gt)*mask|, slots 64..127 to the sum of the SSIM map) followed at byte 1024 by the three
 * partial-derivative maps the backward consumes.  l1_loss = sum(slots 0..63)/(3HW), ssim = sum(slots 64..127)/(3HW); with_ssim = 0 (lambda_dssim == 0) skips the SSIM work, then coeff_ssim must be 0.
 * Backward: dL_dimage = coeff_l1 * d(sum|.|)/dimage + coeff_ssim * d(sum ssim)/dimage
 * (the caller folds 1/(3HW) and lambda_dssim into the two coefficients; `upstream_scalar`, a DEVICE float
 * or NULL, multiplies both -- the loss gradient never visits the host).                                  */
int gg_photometric_workspace_bytes(int32_t width, int32_t height, size_t* map_bytes);
int gg_photometric_forward(int32_t width, int32_t height, const float* image, const float* gt, const float* mask,
                           void* map_ws, int32_t with_ssim, int device, void* stream);
/* L1-only loss (lambda_dssim = 0) against an 8-bit ground truth gt_u8[3,H,W] (= value / 255: frames as they are stored
 * and shipped over PCIe; dequantised on the fly).  dL_dimage == NULL: forward (fills the accumulator slots of map_ws like
 * gg_photometric_forward(with_ssim = 0)); dL_dimage != NULL: backward (like gg_photometric_backward(coeff_ssim = 0)).
 * Needs W*H % 4 == 0.                                                                                               */
int gg_photometric_l1_u8(int32_t width, int32_t height, const float* image, const uint8_t* gt_u8, const float* mask,
                         void* map_ws, float coeff_l1, const float* upstream_scalar, float* dL_dimage, int device,
                         void* stream);
/* folds the accumulator slots gg_photometric_forward left in map_ws into out3 (device float[3]) =
 * (total, l1_loss, ssim) with total = l1_loss (1 - lambda_dssim) + 1 - ssim lambda_dssim
 * (= loss_dict['img'] + loss_dict['ssim'] of s2_registration.py:259-260); nothing visits the host.              */
int gg_photometric_reduce(int32_t width, int32_t height, const void* map_ws, float lambda_dssim, float* out3, int device,
                          void* stream);
int gg_photometric_backward(int32_t width, int32_t height, const float* image, const float* gt, const float* mask,
                            const void* map_ws, float coeff_l1, float coeff_ssim, const float* upstream_scalar,
                            float* dL_dimage, int device, void* stream);

/* ---- introspection ------------------------------------------------------------------------ */
/* copies stage-1 per-Gaussian records out of geom_ws for stage-wise parity tests
 * (xy[N,2], depth[N], conic_opacity[N,4], rgb[N,3], rect[N,2] uint32 packed as
 * x0 | y0<<16, x1 | y1<<16); any may be NULL.                                             */
int gg_debug_read_geom(const gg_view* view, const void* geom_ws, float* xy, float* depth,
                       float* conic_opacity, float* rgb, uint32_t* rect, int device, void* stream);
/* copies the binning result out for index-level parity tests (rows a6-a8): tile_offsets[T+1] (uint32, exclusive
 * scan of the per-tile instance counts; tile_offsets[T] = K) and, for the full-sort forward path, the Gaussian id of
 * every sorted instance sorted_ids[instance_capacity] (uint32; entries of tile t are
 * sorted_ids[tile_offsets[t] .. tile_offsets[t+1]) in front-to-back order).  Either output may be NULL; device or
 * host destinations.                                                                        */
int gg_debug_read_binning(const gg_view* view, const void* tile_ws, const void* record_ws, int64_t instance_capacity,
                          uint32_t* tile_offsets, uint32_t* sorted_ids, int device, void* stream);
/* diagnostics for the dense-scene ("lazy") fused forward, which orders and blends a tile inside ONE kernel: while
 * counters2 (two device uint64, caller-zeroed) is set, every tile adds the SM cycles it spent ordering (depth bucketing
 * + per-bucket sorts) to counters2[0] and packing + blending to counters2[1] -- the honest sort-vs-blend split of that
 * kernel (BASELINE.json configs[4]).  NULL switches it off (default).  Process-wide, not for concurrent use.        */
int gg_debug_lazy_phase_counters(uint64_t* counters2);
/* number of kernels this library launched (process-wide: backward runs on an autograd worker
 * thread) since the last reset.                                                            */
int64_t gg_launch_count(int reset);
/* Optional per-kernel timing (off by default): when enabled every kernel launch is bracketed
 * by CUDA events on the launching stream; gg_kernel_times() waits for the most recent launch
 * of each kernel and writes its duration in ms into ms_out[gg_kernel_count()] (-1 = not
 * launched since enabling).  Meant for bench.py's roofline block, not for the timed region. */
int gg_kernel_timing(int enable);
int gg_kernel_count(void);
const char* gg_kernel_name(int slot);
int gg_kernel_times(float* ms_out);
const char* gg_last_error(void);
const char* gg_version(void);
int gg_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GG_RASTER_H_ */
