/* ptd.h - C ABI of the B200-native path-trace + denoise hot paths.
 *
 * Drop-in boundary for Black-Phoenix/Ai-Path-Tracer-Denoiser's render loop (Inference/src/main.cpp:120-168):
 *   HP-1  pathtraceInit / pathtrace / pathtraceFree           (Inference/src/pathtrace.h:6-8)
 *   HP-2  network_prediction_faster_version(float* rgb)        (Inference/src/main.cpp:101-118),
 *         i.e. torch::jit Module.forward([1,10,H,W]) -> [1,3,H,W] of training/recurrent_autoencoder_model.py
 * plus the scene-file loader the north-star says to keep (Inference/src/scene.cpp:11-320).
 *
 * Conventions: plain C, POD pointers and sizes only; every function returns a ptd_status (0 = ok,
 * negative = error, text via ptd_last_error()); nothing throws or exits across the boundary; handles
 * are opaque and own all device memory; "dev" pointers are CUDA device pointers on the handle's
 * device, `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * There is NO CPU fallback: every compute entry point fails with PTD_ERR_CUDA when no device exists.
 */
#ifndef PTD_H
#define PTD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int ptd_status;
enum {
    PTD_OK = 0,
    PTD_ERR_ARG = -1,      /* bad argument / null handle                       */
    PTD_ERR_IO = -2,       /* file missing or unreadable                        */
    PTD_ERR_PARSE = -3,    /* scene / OBJ / weight file malformed               */
    PTD_ERR_CUDA = -4,     /* CUDA runtime error or no device                   */
    PTD_ERR_UNSUPPORTED = -5,
    PTD_ERR_STATE = -6     /* call order (e.g. dump without trace enabled)      */
};

/* ---- record layouts: byte-identical to Inference/src/sceneStructs.h ------------------------------ */
typedef struct { float x, y, z; } ptd_vec3;
typedef struct { ptd_vec3 origin, direction; } ptd_ray;                                   /* :15-18        */
typedef struct { int type, materialid; ptd_vec3 translation, rotation, scale;
                 float transform[16], inverseTransform[16], invTranspose[16];             /* column major  */
                 ptd_vec3 vel; } ptd_geom;                                                /* :20-30, 248 B */
typedef struct { ptd_vec3 v[3]; ptd_vec3 n[3]; int materialid; } ptd_face;                /* :40-44,  76 B */
typedef struct { ptd_vec3 color; float specular_exponent; ptd_vec3 specular_color;
                 float hasReflective, hasRefractive, indexOfRefraction, emittance; } ptd_material; /* :46-56, 44 B */
typedef struct { int res_x, res_y; ptd_vec3 position, lookAt, view, up, right;
                 float fov_x, fov_y, pixelLength_x, pixelLength_y; } ptd_camera;          /* :58-67,  84 B */
typedef struct { ptd_ray ray; ptd_vec3 color; int pixelIndex, remainingBounces; } ptd_path_segment; /* :77-82, 44 B */
typedef struct { ptd_vec3 lb, ub; } ptd_aabb;                                             /* :84-87,  24 B */
typedef struct { float t; ptd_vec3 surfaceNormal; int materialId; unsigned char is_inside, pad[3];
                 ptd_vec3 intersect; } ptd_intersection;                                  /* :91-97,  36 B */
enum { PTD_SPHERE = 0, PTD_CUBE = 1 };                                                    /* :10-13        */

typedef struct ptd_scene ptd_scene;
typedef struct ptd_pt ptd_pt;
typedef struct ptd_dn ptd_dn;

const char* ptd_last_error(void);          /* thread-local message of the last failing call */
int ptd_version(void);
int ptd_sizeof(int which);                 /* 0 path, 1 intersection, 2 geom, 3 face, 4 material, 5 camera, 6 aabb */
int ptd_device_count(void);                /* 0 when no CUDA device / driver */

/* ---- scene (replaces `new Scene(file)`, scene.cpp:11-42; same text grammar, OBJ mesh ingest) ----- */
ptd_status ptd_scene_load(const char* scene_txt_path, ptd_scene** out);
/* Build a scene from caller arrays (the exact sceneStructs.h records); arrays are copied. */
ptd_status ptd_scene_from_arrays(int ngeoms, const ptd_geom* geoms, int nmaterials, const ptd_material* materials,
                                 int nfaces, const ptd_face* faces, const ptd_aabb* mesh_box, const ptd_camera* camera,
                                 int trace_depth, int iterations, ptd_scene** out);
void ptd_scene_free(ptd_scene*);
ptd_status ptd_scene_counts(const ptd_scene*, int out[5]);   /* geoms, materials, faces, traceDepth, iterations */
const ptd_geom* ptd_scene_geoms(const ptd_scene*);
const ptd_material* ptd_scene_materials(const ptd_scene*);
const ptd_face* ptd_scene_faces(const ptd_scene*);
const ptd_aabb* ptd_scene_mesh_box(const ptd_scene*);
ptd_camera* ptd_scene_camera(ptd_scene*);                    /* mutable, like scene->state.camera */
ptd_status ptd_scene_set_resolution(ptd_scene*, int width, int height);   /* re-derives fov.x / pixelLength (scene.cpp:142-150) */
ptd_status ptd_scene_set_depth(ptd_scene*, int trace_depth);

/* Orbit camera of the frame loop: main.cpp:66-78 (parameters) and main.cpp:126-138 (rebuild). */
ptd_status ptd_camera_orbit_params(const ptd_camera*, float* zoom, float* phi, float* theta);
ptd_status ptd_camera_orbit(ptd_camera*, float zoom, float phi, float theta);

/* ---- HP-1 path tracer (replaces pathtraceInit / pathtrace / pathtraceFree) ----------------------- */
enum {
    PTD_PT_SORT_MATERIAL = 1u,    /* SORT_MATERIAL true  (pathtrace.cu:21, :508-510); default off as in the reference */
    PTD_PT_TRACE = 2u,            /* keep per-bounce PathSegment / ShadeableIntersection arrays for ptd_pt_dump_*     */
    PTD_PT_NO_BVH = 4u,           /* brute-force every face like the reference (pathtrace.cu:258-269); test aid        */
    PTD_PT_KEEP_TERMINATED = 8u,  /* also lay out terminated segments as thrust::partition leaves them (final dump)    */
    PTD_PT_RAY_SORT = 32u,        /* per bounce >= 1, bin the live rays by (Morton cell of the origin, direction octant) and trace them in bin
                                     order: coherent warps for the BVH traversal.  Only the ORDER of tracing changes - every record stays in
                                     its slot, results are bit-identical.  Also switched on by the environment variable PTD_PT_RAY_SORT=1
                                     (PTD_PT_RAY_SORT_BITS = cell bits per axis, 1..5; PTD_PT_RAY_SORT_REFILL = refill threshold; PTD_PT_RAY_SORT_FROM =
                                     first binned bounce, default 2: bounce 1 is still origin-coherent by pixel order)    */
    PTD_PT_GATED_MAIL = 16u       /* row-strip mode: a one-warp gate kernel ahead of every pt_shade waits for the live-count mail of the
                                     strips above, so no shade block ever spins while holding an SM (needed when the path tracer and the
                                     denoiser of a strip run on two streams; see DESIGN.md section 4 "frame loop")            */
};
/* Uploads the scene once (the reference re-uploads every frame, main.cpp:143-146) and builds the BVH. */
ptd_status ptd_pt_create(const ptd_scene*, int device, unsigned flags, ptd_pt** out);
void ptd_pt_destroy(ptd_pt*);
/* Row-strip mode (multi-GPU): the handle traces image rows [row0, row0 + rows) only and writes those rows of the FRAME-sized
 * G-buffer.  Results are bit-identical to the untiled render: the frame-wide compacted index that seeds the reference's RNG
 * (pathtrace.cu:351) is rebuilt on the device from live counts the strips mail each other over NVLink.  Setup: every rank
 * exports a ptd_pt_strip_info_size()-byte blob, the ranks all-gather them (host side) and each passes the concatenation, in
 * strip order, to ptd_pt_strip_connect.  All ranks must render the same camera / iter in the same order.  No material sort. */
ptd_status ptd_pt_create_strip(const ptd_scene*, int device, unsigned flags, int row0, int rows, ptd_pt** out);
int ptd_pt_strip_info_size(void);
ptd_status ptd_pt_strip_export(ptd_pt*, void* info_out, int capacity);
ptd_status ptd_pt_strip_connect(ptd_pt*, const void* infos_in_strip_order, int nranks, int my_rank);
/* Strips living in ONE process (tests; single-process multi-GPU): issues the iteration bounce by bounce across the handles. */
ptd_status ptd_pt_render_group(ptd_pt** strips, int n, const ptd_camera* cam, int iter, float* const* gbuffers_dev, void* const* streams);
/* One 1-spp iteration == pathtrace(pbo, frame, iter) (pathtrace.cu:422-528) for camera `cam` (NULL = the scene's).
 * Writes the 10-plane fp32 G-buffer [10][H][W] (pathtrace.cu:81-94,295-304,379-387; x-mirrored like the
 * reference) to device memory.  iter == 1 starts a new accumulation (what pathtraceInit's memsets do). Asynchronous. */
ptd_status ptd_pt_render(ptd_pt*, const ptd_camera* cam, int iter, float* gbuffer_dev, void* stream);
/* Same + blocking copy into the caller's host_tensor: the reference's pathtrace.cu:525 contract. */
ptd_status ptd_pt_render_host(ptd_pt*, const ptd_camera* cam, int iter, float* host_tensor);
/* RGBA8 view of image/iter (sendImageToPBO, pathtrace.cu:59-79) of the last render, device pointer. */
ptd_status ptd_pt_export_rgba8(ptd_pt*, int iter, unsigned char* pbo_dev, void* stream);
/* Introspection / parity taps (blocking). live_counts: paths entering each bounce (trace_depth ints). */
ptd_status ptd_pt_live_counts(ptd_pt*, int* live_counts, int capacity, int* bounces_run);
ptd_status ptd_pt_dump_paths(ptd_pt*, int bounce, ptd_path_segment* host, int capacity, int* n);          /* needs PTD_PT_TRACE */
ptd_status ptd_pt_dump_intersections(ptd_pt*, int bounce, ptd_intersection* host, int capacity, int* n);  /* needs PTD_PT_TRACE */
ptd_status ptd_pt_dump_final_paths(ptd_pt*, ptd_path_segment* host, int capacity);    /* needs PTD_PT_KEEP_TERMINATED */
ptd_status ptd_pt_dump_image(ptd_pt*, float* host_rgb /* [P][3] */);
ptd_status ptd_pt_bvh_stats(const ptd_pt*, int* nodes, int* leaves, int* max_leaf, int* max_depth);
/* Host-side probe of the BVH ptd_pt_create would build for the scene's mesh (no GPU needed): `nrays` seeded diffuse-bounce-like rays
 * walk the 4-wide layout with pt_trace's traversal rules; the first `brute_rays` of them are checked against the loop over every
 * face (pathtrace.cu:258-269).  out[0] = rays whose (face, t) differ (must be 0), out[1] = interior-node visits per ray,
 * out[2] = triangle tests per ray, out[3] = deepest stack, out[4] = 4-wide nodes, out[5] = leaves, out[6] = mean used children per
 * node, out[7] = fraction of rays that hit. */
ptd_status ptd_bvh_probe(const ptd_scene*, int nrays, unsigned seed, int brute_rays, double out[8]);
/* Host-side estimate of what PTD_PT_RAY_SORT buys: the probe's rays grouped into warps of 32 in arrival order and in bin order
 * (cell_bits per axis + direction octant); out[0], out[1] = L1 wavefronts per ray (one per distinct 128-byte line per load) in
 * arrival / bin order, out[2], out[3] = warp steps per ray, out[4] = bins in use, out[5] = rays. */
ptd_status ptd_bvh_probe_order(const ptd_scene*, int nrays, unsigned seed, int cell_bits, double out[8]);

/* ---- HP-2 recurrent denoising autoencoder (replaces network_prediction_faster_version) ----------- */
enum {
    PTD_DN_FP32 = 0u,             /* fp32 FFMA convolutions (strict-parity path)                                   */
    PTD_DN_TF32 = 1u,             /* tcgen05 kind::tf32 tensor-core convolutions, fp32 accumulate in TMEM (default of the CLI) */
    PTD_DN_3XTF32 = 2u,           /* tcgen05 kind::tf32 with hi/lo operand splitting (hi*hi + hi*lo + lo*hi, the cross terms in their own accumulator) */
    PTD_DN_F16 = 3u,              /* fp16 activation storage (same 10-bit mantissa as tf32, half the HBM / shared-memory bytes) + tcgen05 kind::f16,
                                     fp32 accumulate; the denoised frame itself is written in fp32.  Same stated tolerance as PTD_DN_TF32. */
    PTD_DN_2XF16 = 5u,            /* THE CONTRACT MODE of the tensor-core engine (rel-L2 <= 1e-5, max-abs <= 1e-4 against the fp32 reference model):
                                     every fp32 value v is stored as two fp16 numbers, hi = f16(v) and lo = f16((v - hi) * 2^11) - 22 mantissa bits - and
                                     every conv runs hi*hi + (hi*lo + lo*hi) * 2^-11 on tcgen05 kind::f16 (K = 16 per MMA: half the MMAs of 3xTF32), the
                                     hi*hi chain cut into up to four TMEM accumulators and the cross terms kept in a fifth, summed in fp32 by the
                                     epilogue.  Same bytes per activation as fp32 storage.  |G-buffer values| are clamped to 65504 (fp16 range). */
    PTD_DN_FP32_BATCH_STATS = 4u  /* TorchScript-export compatibility: the fp32 engine with every BatchNorm normalising by the
                                     statistics of its current input, as the module traced by convert_to_torchscript.py:26-30 (never put in
                                     eval mode) does; call with reset_hidden = 1 every frame to mimic its j == 0.  Slow path (4 launches per layer). */
};
/* weights_path: "PTDW" flat dump of the model's state_dict (ai_path_tracer_denoiser_b200/weights.py).
 * H, W: frame size (any; zero-padded bottom/right to a multiple of 32 internally, output cropped). */
ptd_status ptd_dn_create(const char* weights_path, int H, int W, int device, unsigned flags, ptd_dn** out);
void ptd_dn_destroy(ptd_dn*);
/* forward(x, j): gbuffer_dev [10][H][W] fp32 planar -> rgb_dev [3][H][W] fp32 planar.  reset_hidden != 0 is
 * j == 0 (recurrent_autoencoder_model.py:121-128: hidden states zeroed), else the state of the previous call is used. */
ptd_status ptd_dn_forward(ptd_dn*, const float* gbuffer_dev, float* rgb_dev, int reset_hidden, void* stream);
/* Host-pointer form == the reference call site (main.cpp:101-118: H2D of 40*P bytes, forward, D2H of 12*P). Blocking. */
ptd_status ptd_dn_forward_host(ptd_dn*, const float* gbuffer_host, float* rgb_host, int reset_hidden);
/* One frame of runCuda()'s body (main.cpp:143-158) as a single blocking call: path trace with `cam`, denoise on the device (the
 * G-buffer never leaves it), denoised frame [3][H][W] -> rgb_host.  host_tensor (optional, may be NULL) receives the 10-plane
 * G-buffer exactly as pathtrace.cu:525 leaves it in scene->state.host_tensor; its copy overlaps the later bounces and the denoiser.
 * Both handles must live on the same device and cover the whole frame.  Pinned host memory makes the copies asynchronous. */
ptd_status ptd_frame_host(ptd_pt*, ptd_dn*, const ptd_camera* cam, int iter, int reset_hidden, float* host_tensor, float* rgb_host);
/* The same frame asynchronously.  ptd_frame_submit enqueues path trace + denoise + host copies of one frame and returns at once;
 * ptd_frame_wait blocks until the OLDEST submitted frame has reached its host buffers.  Submit frame k + 1 before waiting for frame k
 * and the path trace of k + 1 overlaps the denoiser and the PCIe copies of k.  At most ptd_frame_slots() (3) frames in flight - keep two
 * ahead of the wait and neither the copies nor the host's submission latency ever leave the GPU idle; frames complete in
 * submission order (so the recurrent state is carried in that order); the host buffers of a frame must stay valid, and should be
 * pinned, until its ptd_frame_wait returns.  Results are bit-identical to ptd_frame_host / the two-call path.
 * iter must be 1 (one sample per pixel per frame, what runCuda() renders; PTD_ERR_UNSUPPORTED otherwise).  Both host pointers are optional:
 * with rgb_host == NULL the denoised frame stays on the device (the frame loop itself, nothing copied).
 * Row-strip handles (ptd_pt_create_strip + ptd_dn_create_strip, connected): every rank submits the same frames; host_tensor / rgb_host are
 * still FULL-FRAME [10][H][W] / [3][H][W] buffers of which this rank fills its rows - over all ranks the same bytes reach the host as with
 * one GPU.  A strip's path trace overlaps its denoiser only for handles created with PTD_PT_GATED_MAIL; those also run the two halves on
 * two SM partitions of the GPU (CUDA green contexts: 32 SMs for the denoiser stream, the rest for the path trace; PTD_FRAME_SM_SPLIT=<n>
 * in the environment chooses another size, 0 switches the partition off; it falls back to the shared GPU when the driver has none). */
ptd_status ptd_frame_submit(ptd_pt*, ptd_dn*, const ptd_camera* cam, int iter, int reset_hidden, float* host_tensor, float* rgb_host);
ptd_status ptd_frame_wait(ptd_pt*);
int ptd_frame_slots(void);
/* Device time of a run of submitted frames (CUDA events on the streams ptd_frame_submit launches on): op 0 arms the timer - the next
 * ptd_frame_submit records the start event ahead of its first launch; op 1 records the stop event behind the last submitted frame's
 * denoiser, waits for it and returns the elapsed milliseconds. */
ptd_status ptd_frame_timer(ptd_pt*, int op, float* ms);
/* ---- row-strip mode: the denoiser of ONE frame tiled over several GPUs ----------------------------
 * A strip handle owns padded rows [row0, row0 + rows) (multiples of 32) of the frame.  Its convs store their first / last
 * output row directly into the neighbour strips' halo rows over NVLink (peer pointers) and raise a flag there; the
 * neighbours' next conv waits for the flag on the device.  No host round trip, no NCCL call on the data path.
 * Optionally (environment PTD_DN_REPL_LEVEL=3 on every rank when the strips are created) the levels at 1/8 resolution and below
 * are not tiled: every strip computes them in full from a level-3 input that all strips gather into each other's memory
 * (13 of the 28 convs, 3 % of the FLOPs, no exchange at all).
 * Setup (once): every rank creates its strip, exports a ptd_dn_strip_info_size()-byte POD blob, the ranks all-gather the blobs
 * (torch.distributed on the host side) and each passes the concatenation to ptd_dn_strip_connect.  Needs a tensor-core mode.
 * Per frame every rank calls ptd_dn_forward with the FULL-frame G-buffer [10][H][W] on its own device; it writes rows
 * [row0, min(row0 + rows, H)) of the full-frame rgb [3][H][W].  All ranks must pass the same reset_hidden.
 * Failure detection: a kernel that waits for another strip (halo flag, live-count mail) gives up after 20 s (PTD_SPIN_TIMEOUT_NS
 * at build time) and traps - a dead or mis-ordered peer surfaces as PTD_ERR_CUDA from the next call instead of a hung GPU. */
ptd_status ptd_dn_strip_partition(int H, int nstrips, int index, int* row0, int* rows);   /* even split of the 32-row groups */
ptd_status ptd_dn_create_strip(const char* weights_path, int H, int W, int row0, int rows, int device, unsigned flags, ptd_dn** out);
int ptd_dn_strip_info_size(void);
ptd_status ptd_dn_strip_export(ptd_dn*, void* info_out, int capacity);
ptd_status ptd_dn_strip_connect(ptd_dn*, const void* infos_in_strip_order, int nranks, int my_rank);   /* all strips' blobs, top strip first */
/* Strips living in ONE process (tests; single-process multi-GPU): issues the frame layer by layer across the handles. */
ptd_status ptd_dn_forward_group(ptd_dn** strips, int n, const float* const* gbuffers_dev, float* const* rgbs_dev, int reset_hidden,
                                void* const* streams);
ptd_status ptd_dn_padded_size(const ptd_dn*, int* Hp, int* Wp);
/* Parity tap: copy a hidden state (level 0..5, NCHW fp32, padded size) to host. */
ptd_status ptd_dn_dump_hidden(ptd_dn*, int level, float* host, size_t capacity_floats, int* C, int* H, int* W);
/* Per-launch device timing (cudaEvent pairs on the caller's stream around every kernel of the NEXT forward / render).
 * enable != 0 arms it; the *_times call synchronises and returns milliseconds per launch in launch order
 * (denoiser: pack, 28 convs [+ pools], unpack - names via ptd_dn_launch_name; path tracer: one entry per bounce kernel). */
ptd_status ptd_dn_profile(ptd_dn*, int enable);
ptd_status ptd_dn_launch_times(ptd_dn*, float* ms, int capacity, int* n);
const char* ptd_dn_launch_name(const ptd_dn*, int index);
ptd_status ptd_pt_profile(ptd_pt*, int enable);
ptd_status ptd_pt_launch_times(ptd_pt*, float* ms, int capacity, int* n);
/* Number of kernels one forward launches (for bench.py's gpu_launches). */
int ptd_dn_launches_per_forward(const ptd_dn*);
int ptd_pt_launches_last_render(const ptd_pt*);

#ifdef __cplusplus
}
#endif
#endif /* PTD_H */
