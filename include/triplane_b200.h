/*
 * triplane_b200.h — C ABI of libtriplane_b200.so (sm_100a).
 *
 * B200-native replacement for the differentiable triplane volume-rendering hot path of
 * theEricMa/TriplaneTurbo.  The reference has no C interface for this path: it is PyTorch code
 * calling ATen, the nerfacc CUDA extension and one in-tree pybind module
 * (custom/triplaneturbo/extern/grid_sample_gradfix/gridsample_cuda.cpp:53-56, functions
 * `grad2_2d`/`grad2_3d` taking torch::Tensor).  Every entry point below cites the reference
 * interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types cross this boundary.
 *   - every pointer is a DEVICE pointer on the current CUDA device, owned by the caller
 *     (the Python shim passes torch storage); fp32 unless stated; contiguous.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises, nothing allocates.  Scratch memory is passed in by the caller, sized by
 *     the matching tt_*_floats() query.
 *   - every function returns TT_OK (0) or a negative TT_E_* code; tt_last_error() returns a
 *     thread-local message for the last failure.
 *   - there is NO CPU path: without a CUDA device every compute entry point returns
 *     TT_E_CUDA.
 *
 * Data layout in HBM
 *   - `planes`: channel-last, pre-rotated triplanes  [P][6][R][R][C]  (planes 0-2 geometry,
 *     3-5 texture), produced by tt_repack_planes from the reference's NCHW space cache
 *     [P][6][C][R][R].  One bilinear tap is one contiguous C*4-byte run.
 *   - `wpack`: the three decoder MLPs (64 hidden units, bias-free) in the layouts the kernels
 *     read (row-major and transposed), produced by tt_pack_weights.
 *   - per-ray arrays are [n_rays], per-sample arrays are [n_rays][S] (ray-major, exactly the
 *     reference's flattened `ray_indices` order,
 *     custom/triplaneturbo/models/renderers/generative_space_sdf_volume_renderer.py:317-322).
 */
#ifndef TRIPLANE_B200_H
#define TRIPLANE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_VERSION 100
#define TT_ACC 10 /* per-ray accumulators written by tt_render_fwd */

enum {
    TT_OK = 0,
    TT_E_ARG = -1,      /* bad shape / unsupported size / null pointer */
    TT_E_ALIGN = -2,    /* pointer not 16-byte aligned */
    TT_E_CUDA = -3,     /* CUDA runtime error (message has cudaGetErrorString) */
    TT_E_WORKSPACE = -4 /* scratch too small */
};

/* Scalars of the path (reference: configs/TriplaneTurbo_v1.yaml:73-150). */
typedef struct {
    int32_t C;               /* channels per plane (8,16,32,40,64 compiled in) */
    int32_t R;               /* plane resolution */
    int32_t P;               /* space caches (prompts) */
    int32_t rays_per_cache;  /* consecutive rays that share one space cache (V*H*W) */
    float radius;            /* bbox half-extent (yaml :75) */
    float sdf_bias_radius;   /* sphere SDF bias (yaml :78-79) */
    float inv_std;           /* NeuS 1/std (…sdf_volume_renderer.py:24-35) */
    float cos_anneal_ratio;  /* threestudio/models/renderers/neus_volume_renderer.py:91 */
    float near_plane, far_plane; /* yaml :145-146 */
    float render_step_size;  /* neus_volume_renderer.py:84-86 */
    int32_t flags;           /* TT_FLAG_* */
    int32_t image_h, image_w; /* optional hint: the rays are [B][image_h][image_w] images (row-major pixels).  0 = unknown.
                                 tt_render_bwd then visits the samples in PATCH order (4x4 neighbouring rays x 8 samples per
                                 tile) and merges the taps of a tile on chip before the vector reductions (colour backward
                                 430 -> 379 ms at config 3); results do not depend on it beyond the summation order */
} tt_config;
#define TT_FLAG_PRECISE_BWD 2  /* tt_render_bwd / tt_geometry_bwd: run the colour decoder's backward layers as 3xTF32 (fp32-equivalent)
                                  instead of single-pass TF32; one 128-thread group per CTA fits then (slower) */
#define TT_FLAG_ALL_FEATURES 1 /* tt_render_fwd: evaluate the colour decoder at every sample, not only where the
                                  transmittance is > 0 (needed when `features` is returned to the caller) */

int tt_version(void);
const char* tt_last_error(void);
/* 1 if a CUDA device is usable by this library, else 0 (never falls back to the CPU). */
int tt_device_ok(void);
/* Kernel family: 2 (default) = warp-specialised tcgen05 kernels, 1 = round-1 tcgen05 kernels (3xTF32, TMEM), 0 = SIMT fp32 reference
 * kernels.  Both are CUDA; parity tests run both. */
int tt_set_impl(int impl);
int tt_get_impl(void);
/* Switches of the backward and of the field query; results are identical up to the summation order:
 *   "patch_lists"  1 (default) / 0: with tt_config.image_h/w set, tt_render_bwd visits the samples in 4x4-pixel patch order
 *   "scatter"     -1 auto (default: tile-merged on patch-ordered lists, else run-length merged for rays of >= 256 samples,
 *                  else plain), 0 plain, 1 run-length merged, 2 tile-merged hidden-gradient scatter
 *   "grid_lines"   0 (default) / 1: tt_geometry_fwd on the regular grid gathers z-lines instead of 12 taps per point
 * Initial values come from the environment (TT_SCATTER, TT_PATCH_LISTS). */
int tt_set_option(const char* name, int value);

/* ---- decoder weights -----------------------------------------------------------------
 * Replaces: the nn.Linear parameters of VanillaMLP (threestudio/models/networks.py:67-104)
 * as built at custom/triplaneturbo/models/geometry/few_step_triplane_dual_stable_diffusion.py:101-122.
 * Inputs are nn.Linear.weight tensors ([out,in] row-major): sdf 64xC, 64x64, 1x64;
 * feature 64x3C, 64x64, 3x64; deformation 64xC, 64x64, 3x64 (all three nullable together). */
size_t tt_wpack_floats(int C);
size_t tt_wgrad_floats(int C);
int tt_pack_weights(const float* sdf0, const float* sdf1, const float* sdf2,
                    const float* feat0, const float* feat1, const float* feat2,
                    const float* def0, const float* def1, const float* def2,
                    int C, float* wpack, void* stream);
/* Offsets (in floats) of the [out,in] gradient blocks inside a wgrad buffer:
 * order sdf0, sdf1, sdf2, feat0, feat1, feat2. */
int tt_wgrad_offsets(int C, int64_t offsets[6]);

/* ---- plane repack --------------------------------------------------------------------
 * Replaces: the channel split of `decode` (few_step…diffusion.py:180-196) and the per-call
 * plane rotation of `interpolate_encodings` (few_step…diffusion.py:198-239), done once.
 * src: [P][6][Csrc][R][R]; geometry planes take channels [c_off_geo, c_off_geo+C), texture
 * planes [c_off_tex, c_off_tex+C).  dst: [P][6][R][R][C] rotated (rotate_planes v1). */
int tt_repack_planes(const float* src, int P, int Csrc, int c_off_geo, int c_off_tex, int C, int R,
                     float* dst, void* stream);
/* Adjoint: channel-last rotated gradient -> NCHW [P][6][C][R][R] (overwrites dst). */
int tt_repack_planes_bwd(const float* gplanes, int P, int C, int R, float* gsrc, void* stream);
/* Adjoint of the repack WITH the channel split folded in: the VAE decoder's raw output [P][6][Cdst][R][R] (Cdst = 2C,
 * custom/triplaneturbo/extern/few_step_triplane_dual_sd_modules.py:1012-1020) is the tensor that receives the gradient;
 * geometry planes write channels [c_off_geo, c_off_geo+C), texture planes [c_off_tex, c_off_tex+C); the other channels
 * are left untouched (the caller zero-fills, like the gradient of the reference's masked gather). */
int tt_repack_planes_bwd_split(const float* gplanes, int P, int Cdst, int c_off_geo, int c_off_tex, int C, int R,
                               float* gsrc, void* stream);

/* ---- geometry on point lists -----------------------------------------------------------
 * Replaces: forward / forward_sdf / forward_field / export of the geometry plugin
 * (few_step…diffusion.py:273-430; triplaneturbo_executable/models/geometry/sd_dual_triplanes.py:310-386).
 * points: [P][M][3] world coordinates, or NULL with grid_res>0 for the isosurface grid
 * (threestudio/models/isosurface.py:37-51: index (ix*res+iy)*res+iz, linspace(0,1)->(-1,1)),
 * in which case M = grid_res^3.  Nullable outputs are skipped:
 * sdf,sdf_orig [P*M]; features [P*M][3]; normal,sdf_grad [P*M][3]; deformation [P*M][3]. */
int tt_geometry_fwd(const float* planes, const float* wpack, const tt_config* cfg,
                    const float* points, int64_t M, int grid_res,
                    float* sdf, float* sdf_orig, float* features, float* normal, float* sdf_grad,
                    float* deformation, void* stream);
/* Backward of tt_geometry_fwd w.r.t. planes and decoder weights (second order through the
 * analytic normal; replaces autograd + grid_sample_gradfix/gridsample_cuda.cu:27-210).
 * Upstream gradients (nullable): g_sdf [N] (sum of the grads of sdf and sdf_orig),
 * g_features [N][3], g_normal [N][3], g_sdf_grad [N][3].  scratch: tt_geometry_bwd_scratch_floats(cfg, P*M)
 * floats (per-point seeds, masks and the hidden-gradient planes [P][3][R*R][64] of the colour backward).  gplanes/gw are ACCUMULATED into. */
size_t tt_geometry_bwd_scratch_floats(const tt_config* cfg, int64_t n_points);
int tt_geometry_bwd(const float* planes, const float* wpack, const tt_config* cfg,
                    const float* points, int64_t M,
                    const float* g_sdf, const float* g_features, const float* g_normal,
                    const float* g_sdf_grad, float* scratch, float* gplanes, float* gw, void* stream);

/* Backward of the field query of the mesh paths: forward_field = sdf + deformation decoders on one geometry
 * encoding (few_step…diffusion.py:375-394, trained through at
 * custom/triplaneturbo/models/renderers/generative_space_mesh_rasterize_renderer.py:449-452).
 * g_sdf [N], g_deformation [N][3] (nullable).  gplanes and gw (tt_wgrad_floats, sdf blocks) are ACCUMULATED into;
 * gw_def (tt_wgrad_def_floats: [64][C] | [64][64] | [3][64], nn.Linear layout) likewise.
 * scratch: tt_geometry_bwd_scratch_floats(cfg, P*M) floats. */
size_t tt_wgrad_def_floats(int C);
int tt_field_bwd(const float* planes, const float* wpack, const tt_config* cfg,
                 const float* points, int64_t M, const float* g_sdf, const float* g_deformation,
                 float* scratch, float* gplanes, float* gw, float* gw_def, void* stream);

/* ---- importance sampling ---------------------------------------------------------------
 * Replaces: ImportanceEstimator.sampling (threestudio/models/estimators.py:22-101) with the
 * proposal closure of the renderer (…sdf_volume_renderer.py:243-316) and the two
 * nerfacc.importance_sampling / render_transmittance_from_density calls.
 * rays_o, rays_d: [n_rays][3].  jitter0/jitter1: [n_rays] stratified offsets in [0,1) or NULL
 * (non-stratified, u_j = j/n).  scratch: tt_sample_scratch_floats() floats.
 * t_vals out: [n_rays][n_imp+n_fine+2] sorted interval edges (t_starts = [:, :-1], t_ends = [:, 1:]). */
size_t tt_sample_scratch_floats(int64_t n_rays, int n_imp);
int tt_importance_sample(const float* planes, const float* wpack, const tt_config* cfg,
                         const float* rays_o, const float* rays_d, int64_t n_rays,
                         int n_imp, int n_fine, const float* jitter0, const float* jitter1,
                         float* scratch, float* t_vals, void* stream);

/* ---- stand-alone plane sampler ------------------------------------------------------------
 * Replaces, as an operator of its own: grid_sample / sample_from_planes
 * (custom/triplaneturbo/models/geometry/utils.py:21-24,127-145), the custom autograd pair
 * _GridSample2dForward / _GridSample2dBackward (extern/grid_sample_gradfix/cuda_gridsample.py:22-79) and the
 * second-derivative kernel grid_sampler_2d_grad2_kernel (extern/grid_sample_gradfix/gridsample_cuda.cu:27-210);
 * bilinear, zeros padding, align_corners = False.
 * planes [N*K][H][W][C] channel-last (tt_to_channel_last converts [B][C][HW] NCHW data; 1 <= K <= 4, C % 4 == 0),
 * grid [N*K][M][2] normalised (x -> W, y -> H), out [N][M][OS]: the K planes are summed (concat = 0, OS = C,
 * interpolate_feat "v1") or concatenated (concat = 1, OS = K*C, "v2").
 * bwd: g_planes (same layout as planes) is ACCUMULATED into, g_grid [N*K][M][2] is written; both nullable.
 * bwdbwd(gg_planes, gg_grid = gradients arriving at bwd's two outputs, nullable) -> gg_out [N][M][OS] (written),
 * g_planes (accumulated), g_grid (written); all nullable. */
int tt_to_channel_last(const float* src, int64_t B, int C, int64_t HW, float* dst, void* stream);
int tt_from_channel_last(const float* src, int64_t B, int C, int64_t HW, float* dst, void* stream);
int tt_sample_planes_fwd(const float* planes, int N, int K, int C, int H, int W, const float* grid, int64_t M,
                         int concat, float* out, void* stream);
int tt_sample_planes_bwd(const float* planes, int N, int K, int C, int H, int W, const float* grid, int64_t M,
                         int concat, const float* g_out, float* g_planes, float* g_grid, void* stream);
int tt_sample_planes_bwdbwd(const float* planes, int N, int K, int C, int H, int W, const float* grid, int64_t M,
                            int concat, const float* g_out, const float* gg_planes, const float* gg_grid,
                            float* gg_out, float* g_planes, float* g_grid, void* stream);

/* ---- fused march + decoder MLPs + NeuS alpha + compositing -------------------------------
 * Replaces: GenerativeSpaceSDFVolumeRenderer._forward from sample positions to the per-ray
 * accumulators (…sdf_volume_renderer.py:317-431,466-472), i.e. geometry.forward with analytic
 * normals, NoMaterial, NeuSVolumeRenderer.get_alpha (neus_volume_renderer.py:93-117),
 * nerfacc.render_weight_from_alpha and the five nerfacc.accumulate_along_rays calls.
 * t_starts/t_ends: rows of S floats, consecutive rows t_stride floats apart.
 * acc out: [n_rays][TT_ACC] = opacity, depth, rgb[3], z_variance, normal_sum[3] (un-normalised),
 * eikonal_sum = Σ_samples (|sdf_grad| - 1)^2 (the eikonal loss of
 * custom/triplaneturbo/systems/multiprompt_dual_renderer_multistep_generator.py:690-714 without materialising
 * sdf_grad).
 * Per-sample outputs [n_rays*S] (all nullable; needed by tt_render_bwd: sdf, sdf_grad,
 * features, trans): sdf, sdf_orig, sdf_grad[3], normal[3], features[3], weights, trans.
 * masks: [n_rays*S][4] (nullable) 64-bit ReLU masks of the hidden layers: colour decoder [0..1], SDF decoder [2..3]
 *   (32 B/sample), consumed by tt_render_bwd so that it recomputes neither decoder's activations pattern; when
 *   NULL the backward runs the fp32 SIMT kernels.
 * scratch: tt_render_fwd_scratch_floats() floats (holds the per-sample state the caller did not ask for). */
size_t tt_render_fwd_scratch_floats(int64_t n_rays, int S);
int tt_render_fwd(const float* planes, const float* wpack, const tt_config* cfg,
                  const float* rays_o, const float* rays_d, int64_t n_rays,
                  const float* t_starts, const float* t_ends, int64_t t_stride, int S,
                  float* acc,
                  float* sdf, float* sdf_orig, float* sdf_grad, float* normal, float* features,
                  float* weights, float* trans, uint64_t* masks, float* scratch, void* stream);
/* Backward.  g_acc: [n_rays][TT_ACC] gradients of the accumulators.  Per-sample upstream
 * gradients (nullable): g_sdf [N], g_sdf_grad [N][3], g_normal [N][3], g_features [N][3],
 * g_weights [N].  rgb_grad_scale multiplies the colour gradient (rgb_grad_shrink,
 * …sdf_volume_renderer.py:397-400).  scratch: tt_render_bwd_scratch_floats(cfg, n_rays, S) floats
 * (per-sample seeds, sample lists, hidden-gradient planes [P][3][R*R][64] of the colour backward).
 * gplanes [P][6][R][R][C] and gw (tt_wgrad_floats) are ACCUMULATED into; g_inv_std (1 float,
 * nullable) likewise. */
size_t tt_render_bwd_scratch_floats(const tt_config* cfg, int64_t n_rays, int S);
int tt_render_bwd(const float* planes, const float* wpack, const tt_config* cfg,
                  const float* rays_o, const float* rays_d, int64_t n_rays,
                  const float* t_starts, const float* t_ends, int64_t t_stride, int S,
                  const float* acc, const float* sdf, const float* sdf_grad, const float* features,
                  const float* trans, const uint64_t* masks,
                  const float* g_acc, const float* g_sdf, const float* g_sdf_grad,
                  const float* g_normal, const float* g_features, const float* g_weights,
                  float rgb_grad_scale, float* scratch,
                  float* gplanes, float* gw, float* g_inv_std, void* stream);

/* ---- stand-alone compositor ---------------------------------------------------------------
 * Replaces: nerfacc.render_weight_from_alpha + nerfacc.accumulate_along_rays for the dense
 * ray_indices of the path (…sdf_volume_renderer.py:408-431).  alphas: [n_rays][S];
 * values: [n_rays][S][D] or NULL (D=0); weights,trans: [n_rays][S]; out: [n_rays][max(D,1)]. */
int tt_composite_fwd(const float* alphas, const float* values, int64_t n_rays, int S, int D,
                     float* weights, float* trans, float* out, void* stream);
int tt_composite_bwd(const float* alphas, const float* values, const float* trans,
                     const float* g_out, const float* g_weights, int64_t n_rays, int S, int D,
                     float* g_alphas, float* g_values, void* stream);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t tt_launch_count(void);
/* Per-launch device timing for bench.py's roofline: between begin and end every kernel launch is bracketed by
 * CUDA events on its stream; end synchronises and writes "kernel:milliseconds;" records into buf. */
int tt_profile_begin(void);
int tt_profile_end(char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* TRIPLANE_B200_H */
