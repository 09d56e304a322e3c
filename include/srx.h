/* srx.h — C ABI of libsrx.so: B200 (sm_100a) kernels for Stable-Renderer's correspondence-map latent overlap step
 * and UV-texture bake.
 *
 * The reference (92MING/Stable-Renderer) has no FFI for this path: it is pure Python calling torch ops
 * (SURVEY.md §8b).  Each entry point below therefore names the reference *Python* interface whose arithmetic it
 * replaces (paths relative to the reference root).  The Python wrappers in stable-renderer_b200/ keep those
 * signatures and call these functions through ctypes; INTEGRATION.md shows the binding a reference maintainer
 * would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative srx_status; srx_last_error() returns the message of the
 *     last failure on the calling thread.  No exceptions cross the boundary.
 *   - all `*_dev` / `void*` data pointers are DEVICE pointers owned by the caller (torch tensors); the library
 *     never frees them.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - nothing synchronises the host unless stated ("syncs").
 *   - handles are opaque and destroyed explicitly.
 */
#ifndef SRX_H_
#define SRX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRX_VERSION 100 /* 0.1.0 */

typedef enum srx_status {
    SRX_OK = 0,
    SRX_ERR_INVALID = -1,     /* bad argument (maps to ValueError) */
    SRX_ERR_INDEX = -2,       /* an entry addresses a latent cell / texel out of range (maps to IndexError) */
    SRX_ERR_CUDA = -3,        /* CUDA runtime error */
    SRX_ERR_UNSUPPORTED = -4, /* valid request that this build does not implement */
    SRX_ERR_KEY_RANGE = -5,   /* key outside the dense slot table */
    SRX_ERR_PEER_LOST = -6    /* a bounded wait inside the frame-sharded step kernel timed out (srx_plan_check) */
} srx_status;

typedef enum srx_dtype {
    SRX_F32 = 0, SRX_F16 = 1, SRX_BF16 = 2, /* latents / colours */
    SRX_I32 = 10, SRX_I16 = 11,             /* id buffers (RGBA_32I attachment, or the int16 .npy dumps) */
    SRX_U8 = 20
} srx_dtype;

/* What identifies "the same surface point" across frames. */
typedef enum srx_key_mode {
    /* current generation: float32(vertexID) only — corresponder.py:331-334 */
    SRX_KEY_VERTEX = 0,
    /* legacy generation: the whole id 4-tuple, optionally merged (texX//d, texY//d) —
       legacy_codes/stable_rendering_algo/data_classes/correspondence_map.py:153-168, 276-286 */
    SRX_KEY_TUPLE = 1
} srx_key_mode;

typedef enum srx_strategy { /* legacy_codes/stable_rendering_algo/overlap/algorithms.py:34-133 */
    SRX_STRATEGY_AVERAGE = 0,
    SRX_STRATEGY_FRAME_DISTANCE = 1,
    SRX_STRATEGY_PIXEL_DISTANCE = 2,
    SRX_STRATEGY_VIEW_NORMAL = 3
} srx_strategy;

typedef enum srx_accum_mode {
    SRX_ACCUM_FAST = 0,         /* float32 vector atomics (order not fixed; within 1e-5 of the reference) */
    SRX_ACCUM_DETERMINISTIC = 1,/* Q31.32 fixed-point int64 atomics: order independent, bit reproducible */
    SRX_ACCUM_FAST_SPLIT = 2    /* as FAST, but srx_overlap_step always runs the split reduce + gather kernels instead of
                                   the single persistent kernel (profiling / cross-checking) */
} srx_accum_mode;

int srx_version(void);
const char *srx_last_error(void);
/* number of SMs of the current device (grid sizing is a multiple of this) */
int srx_device_sm_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Keying plan  — replaces IDMap.create_vertex_screen_info (source/engine/static/corrmap.py:220-280) and the
 * coordinate step of OverlapCorresponder.step_finished (source/common_utils/stable_render_utils/corresponder.py:310-314)
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct srx_plan srx_plan;

typedef struct srx_plan_desc {
    int id_dtype;            /* SRX_I32 | SRX_I16 */
    int frames;              /* F: id frames */
    int height, width;       /* H, W of the id buffers */
    int batch;               /* B: latent frames */
    int channels;            /* C: latent channels (4 for SD1.5 / SDXL) */
    int lat_h, lat_w;        /* h, w of the latents */
    int key_mode;            /* srx_key_mode */
    int merge_len;           /* SRX_KEY_TUPLE only: merge_nearby distance, 0/1 = off */
    int accum_mode;          /* srx_accum_mode */
    int64_t key_capacity;    /* dense slot count; 0 = derive with a scan of the ids (syncs) */
    const int32_t *frame_map;/* HOST array [F]: value used as batch index of each id frame (corrmap.py:251-253,
                                corresponder.py:314); NULL = identity.  Negative values wrap like torch indexing. */
} srx_plan_desc;

typedef struct srx_plan_info {
    int64_t n_valid;         /* entries (valid pixels); -1 when no scan ran */
    int64_t key_min, key_max;/* over valid pixels; meaningful when a scan ran */
    int64_t key_capacity;    /* slots in the dense accumulator */
    int64_t workspace_bytes; /* device bytes the caller must bind with srx_plan_bind_workspace */
    int64_t accum_offset;    /* byte offset / size of the accumulator block inside the workspace ... */
    int64_t accum_bytes;     /* ... i.e. the buffer a multi-GPU caller all-reduces between reduce and gather */
    int accum_dtype;         /* SRX_F32 (fast) — int64 in deterministic mode is reported as -64 */
    int fast_path;           /* 1 when the 8x8-pixels-per-cell warp kernel applies */
    int fused;               /* 1 when srx_overlap_step runs as the single persistent kernel (and peer mode is available) */
    int64_t need_offset;     /* byte offset inside the workspace of the [key_capacity] byte map "key wins a cell", written by
                                the first call of srx_plan_build_cache; frame-sharded callers MAX-all-reduce it across ranks
                                before the second call (a key matters if it wins a cell on ANY rank) */
} srx_plan_info;

/* Builds the plan for one batch of id buffers.  `ids_dev` may be NULL when key_capacity > 0 (ids are then given to
 * each call).  Syncs only when key_capacity == 0 (scan).  Errors: SRX_ERR_INDEX when a valid pixel maps outside
 * the latent (the reference raises IndexError at corresponder.py:324-329). */
int srx_plan_create(srx_plan **out, const srx_plan_desc *desc, const void *ids_dev, void *stream);
int srx_plan_get_info(const srx_plan *plan, srx_plan_info *info);
int srx_plan_bind_workspace(srx_plan *plan, void *workspace_dev, int64_t bytes, void *stream);
/* Frame-sharded peer mode: `peer_workspaces[i]` is rank i's bound workspace as mapped into this
 * process (CUDA IPC / symmetric memory; entry `rank` is ignored).  All ranks bind workspaces of identical layout
 * (same key capacity and channels) on GPUs with the same SM count.  Afterwards srx_overlap_step reduces this rank's
 * frames, exchanges the key accumulator with the peers over NVLink inside the same kernel, and gathers this rank's
 * frames; every rank must call it the same number of times.  world = 1 unbinds. */
int srx_plan_bind_peers(srx_plan *plan, int rank, int world, void *const *peer_workspaces);
/* NVLS form of the in-kernel exchange: `mc_ws` = the multicast address of the symmetric workspace (one pointer that reaches
 * every rank's copy through the NVSwitch).  The owner then receives its slice reduced inside the switch (multimem.ld_reduce)
 * and broadcasts the totals (multimem.st); every rank gathers from its LOCAL record table.  All ranks or none; after
 * srx_plan_bind_peers; NULL = back to the pull exchange. */
int srx_plan_bind_multicast(srx_plan *plan, void *mc_ws);
/* Bucketing pass for the cached-plan regime (steps 2..N of a sampling run reuse the ids).  Replaces what the
 * reference recomputes every step — `unique(return_inverse)` over the [N] key column (math_utils.py:137) — by one pass
 * per id batch.  Call twice: with pool_dev == NULL it streams the ids once (per-cell winners, the set of keys that win a
 * cell, pairs per CTA), SYNCS, and returns in *pool_bytes the pool size to allocate; with pool_dev it streams the ids again
 * and stores the (key, cell, multiplicity) pairs of the keys that matter.  The pool must stay alive while cached steps run.
 * Frame-sharded runs combine the ranks' key maps between the two calls (srx_plan_info.need_offset). */
int srx_plan_build_cache(srx_plan *plan, const void *ids_dev, void *pool_dev, int64_t *pool_bytes, void *stream);
/* pairs kept / pairs seen by the last build (syncs) */
int srx_plan_cache_entries(srx_plan *plan, int64_t *kept, int64_t *capacity, void *stream);
/* CTAs of the persistent step kernel (default 0 = one per SM).  Every rank of a peer group must use the same value;
 * smaller grids let several ranks' kernels share one GPU (single-GPU emulation of a frame-sharded run in the tests). */
int srx_plan_set_grid(srx_plan *plan, int ctas);
/* Profiling aid: SM-clock timestamps of the phases of the last persistent step (first CTA in out16[0..7], last CTA in
 * out16[8..15]; index meaning in csrc/srx_fused.cu).  Syncs. */
int srx_plan_read_trace(srx_plan *plan, int64_t *out16, void *stream);
/* Profiling aid: (earliest CTA start, latest CTA end) of the last 32 persistent steps on the GPU global timer [ns],
 * out64[2 * (step & 31) + {0,1}], then the first CTA's phase stamps of those steps, out64[64 + 8 * (step & 31) + idx]
 * (320 values in all).  Syncs. */
int srx_plan_read_step_ring(srx_plan *plan, uint64_t *out64, void *stream);
int srx_plan_destroy(srx_plan *plan);
/* Lazily reported device-side failures of earlier launches (key out of the slot table, cell out of range).
 * Syncs the stream. */
int srx_plan_check(srx_plan *plan, void *stream);

/* [N,7] float32 rows (sprite, material, map_index, vertexID, x/H, y/W, frame value) in (frame,y,x) order —
 * bit-identical to IDMap.create_vertex_screen_info (corrmap.py:220-280).  `out_dev` must hold F*H*W*7 floats
 * (worst case); *n_rows receives N.  Syncs. */
int srx_vertex_screen_info(const void *ids_dev, int id_dtype, int frames, int height, int width,
                           const int32_t *frame_values_host, float *out_dev, int64_t *n_rows, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Overlap step — replaces OverlapCorresponder.step_finished (corresponder.py:298-376):
 * gather, tensor_group_by_then_average (source/common_utils/math_utils.py:86-161), blend, duplicate-index
 * write-back (last entry wins) and adaptive_instance_normalization (math_utils.py:55-80); in place on `x_dev`.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct srx_step_args {
    void *x_dev;             /* latents [B,C,h,w], contiguous, updated in place */
    int x_dtype;             /* SRX_F32 | SRX_F16 | SRX_BF16 */
    const void *ids_dev;     /* id buffers [F,H,W,4] of this step (streaming regime); NULL = cached-plan regime: run from the
                                pairs stored by srx_plan_build_cache (same ids for every denoise step of a sampling run) */
    float ratio;             /* step_finished_inject_ratio (corresponder.py:180,351-352) */
    int adain;               /* 1 = reference behaviour (re-standardise the original latents); 0 = write the blend */
    int cache_slots;         /* reserved, must be 0 */
} srx_step_args;

/* reduce + gather on one GPU */
int srx_overlap_step(srx_plan *plan, const srx_step_args *args, void *stream);
/* split form for frame-sharded multi-GPU runs: reduce into the plan's accumulator, all-reduce that block with NCCL,
 * then gather */
int srx_accum_reduce(srx_plan *plan, const srx_step_args *args, void *stream);
int srx_accum_finalize_gather(srx_plan *plan, const srx_step_args *args, void *stream);

/* tensor_group_by_then_average on explicit columns (math_utils.py:86-161): values [N,C] float32 and float32 keys [N];
 * writes the per-row group mean [N,C].  Keys must be non-negative integers < key_capacity. Workspace: key_capacity*(C+1) floats. */
int srx_group_by_then_average(const float *values_dev, const float *keys_dev, int64_t n, int channels,
                              float *out_dev, float *workspace_dev, int64_t key_capacity, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Same-key broadcast initialisers — replace tensor_group_by_then_randn_init (source/common_utils/math_utils.py:164-229)
 * and the arithmetic of CreateNoiseSequenceFromIdMap (source/comfyUI/stable_rendering/_nodes/loaders.py:193-271).
 * The reference's `unique(return_inverse=True)` sort becomes a dense presence table + exclusive scan; the random rows
 * are still drawn by the caller with torch.randn (same call as the reference => same values on the same device).
 * ------------------------------------------------------------------------------------------------------------------ */
/* ints the rank table workspace must hold for a key capacity */
int64_t srx_group_rank_workspace_ints(int64_t key_capacity);
/* keys [n] float32 (non-negative integers < key_capacity).  On return table_ws[k] = rank of key k among the sorted unique
 * keys or -1, rank_out[i] (optional) = rank of keys[i] = torch.unique's `inverse`, *n_unique = number of unique keys.  Syncs. */
int srx_group_rank(const float *keys_dev, int64_t n, int64_t key_capacity, int32_t *table_ws, int32_t *rank_out,
                   int64_t *n_unique, void *stream);
/* the same table straight from id buffers: key = float32(vertexID) of every pixel kept by corrmap.py:266-275.  Syncs. */
int srx_ids_rank_table(const void *ids_dev, int id_dtype, int frames, int height, int width, int64_t key_capacity,
                       int32_t *table_ws, int64_t *n_unique, void *stream);
/* out[i, :] = table[rank[i], :]  —  random_values[inverse_indices] (math_utils.py:224) */
int srx_group_broadcast(const float *table_dev, const int32_t *rank_dev, int64_t n, int channels, float *out_dev, void *stream);

typedef struct srx_noise_args {
    const void *ids_dev;          /* [F,H,W,4] */
    int id_dtype;
    int frames, height, width;
    const int32_t *inv_frame_dev; /* [F] LAST id frame that writes latent frame f, or -1 (inverse of IDMap.frame_indices) */
    const int32_t *rank_table_dev;/* from srx_ids_rank_table */
    const float *key_latent_dev;  /* [n_unique,4] random row per key for the latent (mode 0 only) */
    const float *key_noise_dev;   /* [n_unique,4] random row per key for the noise */
    const float *base_latent_dev; /* [4,S,S] the node's base latent draw (mode 0 only) */
    const float *base_noise_dev;  /* [4,S,S] */
    float *latent_out_dev;        /* mode 0: [F,4,S/8,S/8] */
    float *noise_out_dev;         /* mode 0: [F,4,S/8,S/8]; modes 1-3: F*4*S*S/32 floats (the node views them as [2F,4,S/8,S/8]) */
    int mode;                     /* downsample_option: 0 nearest, 1 mean, 2 max, 3 min (loaders.py:252-268) */
    int work_size;                /* S = the node's working size (512 SD15 / 1024 SDXL); 0 = the id map's own (square) size.  Entry
                                     (x, y) lands in pixel (trunc(fl32(x/H)*S), trunc(fl32(y/W)*S)), loaders.py:218-219 */
    const int32_t *prev_frame_dev;/* [F] the previous id frame writing the same latent frame as id frame g, or -1; NULL = none */
} srx_noise_args;
int srx_noise_from_ids(const srx_noise_args *args, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Legacy overlap — replaces ResizeOverlap.__call__ / Overlap.__call__ with kernel_radius 0
 * (legacy_codes/stable_rendering_algo/overlap/overlap.py:83-152,180-222) for the four OverlapAlgorithm strategies
 * (overlap/algorithms.py:34-118).  Latents [T,B,C,h,w] (B folded into channels), ids [T,H,W,4] keyed by the whole id
 * 4-tuple like CorrespondenceMap.FromExisting (legacy data_classes/correspondence_map.py:145-170).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct srx_legacy_desc {
    int id_dtype;              /* SRX_I32 | SRX_I16 */
    int frames, height, width; /* T, H, W of the id buffers (= CorrespondenceMap.num_frames / .size) */
    int channels;              /* B*C of one frame's latent [B,C,h,w] */
    int lat_h, lat_w;          /* h, w; equal to H, W for Overlap.__call__, smaller for ResizeOverlap.__call__ */
    int merge_len;             /* merge_nearby distance (correspondence_map.py:276-286); 0/1 = off */
    int strategy;              /* srx_strategy */
} srx_legacy_desc;

typedef struct srx_legacy_args {
    void *x_dev;               /* [T, B*C, h, w] updated in place */
    int x_dtype;               /* SRX_F32 | SRX_F16 | SRX_BF16 */
    const void *ids_dev;       /* [T,H,W,4] */
    float alpha;               /* alpha scheduler value (overlap.py:100,145) */
    const float *view_normal_dev; /* [T,H,W] float32 for SRX_STRATEGY_VIEW_NORMAL (utils.py:56-102), else NULL */
    void *workspace_dev;       /* srx_legacy_workspace_bytes() bytes */
    int64_t workspace_bytes;
    int defer_status;          /* 1 = no host sync inside srx_legacy_overlap: sticky status, reported by srx_legacy_check */
} srx_legacy_args;
int64_t srx_legacy_workspace_bytes(const srx_legacy_desc *desc);
/* Errors (after a sync) with SRX_ERR_KEY_RANGE when an id component does not fit the packed 64-bit key. */
int srx_legacy_overlap(const srx_legacy_desc *desc, const srx_legacy_args *args, void *stream);
int srx_legacy_check(const srx_legacy_desc *desc, const srx_legacy_args *args, void *stream);

/* kernel_radius > 0 (overlap.py:61-80,137-145): the reference pools the diagonal neighbours (y+d, x+d), d in [-r, r], and
 * writes every blended trace into the storage it keeps reading, trace after trace in dict order — a Gauss-Seidel sweep whose
 * result depends on that order.  This entry reproduces it: traces ranked in insertion order, entries in (frame,row,col) order,
 * one CTA walks the traces strictly in sequence (entries of a trace in parallel).  Same descriptor / arguments as
 * srx_legacy_overlap (radius 0 gives the same result as srx_legacy_overlap, more slowly); its own, larger workspace. */
int64_t srx_legacy_ordered_workspace_bytes(const srx_legacy_desc *desc);
int srx_legacy_overlap_ordered(const srx_legacy_desc *desc, const srx_legacy_args *args, int kernel_radius, void *stream);

/* johnny_overlap (legacy_codes/legacy_diffuser/modules/diffuser_pipelines/overlap/johnny_overlap.py:15-141), the diffusers
 * pipeline's experimental variant: frame-distance weights 1/(|dt|+1), every entry of a trace updated IN PLACE before the next
 * one is evaluated, optional mix with a base colour taken at the trace's first entry (`beta`; base_dev = [T, B*C, h, w] float32
 * noised original latents), nearest up- / down-sampling without the `where`.  desc->strategy is ignored; workspace =
 * srx_legacy_ordered_workspace_bytes(). */
int srx_johnny_overlap(const srx_legacy_desc *desc, const srx_legacy_args *args, float beta, const float *base_dev, void *stream);

/* CorrespondenceMap maintenance (legacy data_classes/correspondence_map.py).  The map IS the id buffers: a key is the id
 * tuple (after merge_nearby's floor division), deleting a key clears the ids of every pixel that carries it.
 *   srx_corrmap_first_appearance: mask[i] = 1 where pixel i introduces its key — the marked pixels, in index order, are the
 *       reference dict's keys in insertion order (:148-168); needed to replay dropout_index's random stream (:207-223).
 *   srx_corrmap_drop_keys: delete the keys carried by the seed pixels (dropout_index :219-223, dropout_in_rectangle :268-274).
 * Workspace: srx_corrmap_keys_workspace_bytes(F*H*W) for the first, (n_seeds) for the second.  Both sync. */
int64_t srx_corrmap_keys_workspace_bytes(int64_t n_keys_upper_bound);
int srx_corrmap_first_appearance(const void *ids_dev, int id_dtype, int frames, int height, int width, int merge_len,
                                 uint8_t *mask_out_dev, void *workspace_dev, int64_t workspace_bytes, void *stream);
int srx_corrmap_drop_keys(void *ids_dev, int id_dtype, int frames, int height, int width, int merge_len,
                          const int64_t *seed_pixels_dev, int64_t n_seeds, void *workspace_dev, int64_t workspace_bytes, void *stream);

/* CorrMapLatentNoiseInitializer (legacy_codes/nodes/latent.py:10-40): keys with >= 2 entries ("traces") get one random
 * 4-vector for the latent and one for the noise, drawn in dict insertion order (:28-35), then a nearest down-sample (:37-38).
 *   srx_corrmap_trace_ranks: rank_out[i] = index of pixel i's key among the traces in insertion order, -1 otherwise;
 *       *n_traces_out (host) = number of traces.  Syncs.
 *   srx_corrmap_noise_fill: base [2,4,H,W] (latent base, noise base: one frame each, repeated over the batch, :22-26),
 *       rows [n_traces,2,4] (latent row, noise row) -> latent_out / noise_out [batch,4,lat_h,lat_w], evaluated only at the
 *       pixels the down-sample keeps.  batch < frames is an error (IndexError in the reference). */
int64_t srx_corrmap_trace_ranks_workspace_bytes(int64_t n_pixels);
int srx_corrmap_trace_ranks(const void *ids_dev, int id_dtype, int frames, int height, int width, int merge_len,
                            int32_t *rank_out_dev, int64_t *n_traces_out, void *workspace_dev, int64_t workspace_bytes, void *stream);
int srx_corrmap_noise_fill(const int32_t *rank_dev, int frames, int height, int width, int batch, int lat_h, int lat_w,
                           const float *base_dev, const float *rows_dev, float *latent_out_dev, float *noise_out_dev, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Bake — replaces CorrespondMap.update/_update (source/engine/static/corrmap.py:578-736)
 * ------------------------------------------------------------------------------------------------------------------ */
typedef enum srx_bake_mode { SRX_BAKE_REPLACE = 0, SRX_BAKE_REPLACE_AVG = 1, SRX_BAKE_FIRST = 2, SRX_BAKE_FIRST_AVG = 3 } srx_bake_mode;
typedef enum srx_bake_weight { SRX_WEIGHT_NONE = 0, /* reference behaviour: last pixel of the chosen frame wins */
                               SRX_WEIGHT_UNIFORM = 1, SRX_WEIGHT_VIEW_NORMAL = 2, SRX_WEIGHT_VIEW_NORMAL_DEPTH = 3 } srx_bake_weight;

typedef struct srx_bake_args {
    void *values_dev;          /* fp16 [k2, Ht*Wt, C] atlas (corrmap.py:410), updated in place */
    uint8_t *writtens_dev;     /* bool [k2, Ht*Wt] (corrmap.py:411) */
    int k2, texels, channels;  /* k*k, Ht*Wt, C */
    const void *colors_dev;    /* [F,H,W,Cin] */
    int color_dtype;           /* SRX_F32 | SRX_F16 | SRX_BF16 */
    int color_channels;        /* Cin (3 with C == 4 appends alpha = 1, corrmap.py:683-684; Cin > C truncates, :681-682) */
    const void *ids_dev;       /* [F,H,W,4] */
    int id_dtype;
    const float *masks_dev;    /* [F,H,W] float32 or NULL; pixels with mask > 0 are kept (after inversion) */
    int inverse_masks;         /* corrmap.py:651-654 */
    int frames, height, width;
    int sprite_id, material_id;/* -1 = no filter (None) */
    int ignore_obj_mat_id;
    int mode;                  /* srx_bake_mode */
    int weight_mode;           /* srx_bake_weight; != NONE selects the weighted multi-view bake */
    const void *normal_depth_dev; /* fp16 [F,H,W,4] for the VIEW_NORMAL* weights */
    void *workspace_dev;       /* srx_bake_workspace_bytes() bytes */
    int64_t workspace_bytes;
    int phase;                 /* weighted bake only: 0 = accumulate + finalize (default); 1 = clear + accumulate the weighted
                                  sums of these views into the workspace; 2 = finalize the atlas from the workspace.  View-sharded
                                  multi-GPU bakes run phase 1 per rank, sum the first srx_bake_workspace_bytes() - 256 bytes of
                                  the workspaces as float32 (all-reduce), then phase 2 on every rank.
                                  Reference modes (weight NONE), view-sharded: 1 = claim with order keys that number the views of
                                  all ranks (frame_offset / frames_global below) -> MAX all-reduce of the first k2*texels int32
                                  words of the workspace; 2 = write the texels this rank's views won into the workspace's partial
                                  atlas -> SUM all-reduce of that region as int32 words (each texel is non-zero on one rank only);
                                  3 = copy the claimed texels into the atlas.  Workspace: srx_bake_sharded_workspace_bytes(). */
    int frame_offset;          /* view-sharded reference modes: index of this rank's first view among all ranks' views */
    int frames_global;         /* ... and the number of views of all ranks together (frames_global * H * W < 2^31) */
    int defer_status;          /* 1 = no host sync inside the call: the status word stays sticky in the (initially zeroed)
                                  workspace until srx_bake_check reads it — graph-capturable, like srx_plan_check for the step */
} srx_bake_args;
int64_t srx_bake_workspace_bytes(int k2, int texels, int channels, int weight_mode);
/* [owner words k2*texels*4, 256-aligned][256 status][partial atlas k2*texels*channels*2, 256-aligned] */
int64_t srx_bake_sharded_workspace_bytes(int k2, int texels, int channels);
/* Errors with SRX_ERR_INDEX (after a sync) when a kept pixel addresses a texel outside the atlas. */
int srx_bake_update(const srx_bake_args *args, void *stream);
/* Deferred form (args->defer_status): reports and clears the status of the calls since the last check.  Syncs. */
int srx_bake_check(const srx_bake_args *args, void *stream);

/* On-disk atlas format — the arithmetic of CorrespondMap.dump / Load (source/engine/static/corrmap.py:776-791, 846-858):
 * uint8 = clip(255 * value, 0, 255) evaluated in float16 like numpy does, flags 0 / 255; and back: float32(u8) / 255 stored
 * as float16, flag != 0.  The PNG / meta.json / zip handling stays in Python. */
int srx_atlas_quantize(const void *values_f16_dev, const uint8_t *writtens_dev, uint8_t *out_values_dev, uint8_t *out_flags_dev,
                       int64_t n_values, int64_t n_flags, void *stream);
int srx_atlas_dequantize(const uint8_t *in_values_dev, const uint8_t *in_flags_dev, void *values_f16_dev, uint8_t *writtens_dev,
                         int64_t n_values, int64_t n_flags, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Texture <-> tensor interop — replaces Texture._init_tensor/tensor/set_data (source/engine/static/texture/texture.py:166-254,
 * 326-408: pycuda RegisteredImage + Memcpy2D + torch.cuda.synchronize) and the cuda-python wrappers of
 * source/common_utils/cuda_utils.py:101-190.  The mapped cudaArray is read/written by kernels through a surface
 * object on the caller's stream: no staging copy, no device-wide sync.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct srx_gl_resource srx_gl_resource;
int srx_gl_register_image(srx_gl_resource **out, unsigned int gl_texture, unsigned int gl_target, unsigned int flags);
int srx_gl_map(srx_gl_resource *res, void **cuda_array_out, void *stream);
int srx_gl_unmap(srx_gl_resource *res, void *stream);
int srx_gl_unregister(srx_gl_resource *res);
/* array <-> linear tensor with the GL bottom-left -> top-left row flip (texture.py:236,253) fused in.
 * `cuda_array` is a cudaArray_t (from srx_gl_map, or cudaMallocArray in tests); texel_bytes = channels * bytes per channel. */
int srx_array_to_tensor(void *cuda_array, void *dst_dev, int width, int height, int texel_bytes, int flip, void *stream);
int srx_tensor_to_array(void *cuda_array, const void *src_dev, int width, int height, int texel_bytes, int flip,
                        int x_offset, int y_offset, void *stream);
/* Channel-aware forms for the Texture wrapper (texture.py:221-254, 326-408): the array's own channel count decides how texels
 * are read / written — a three-channel GL texture is a four-channel CUDA array.  array -> tensor [height,width,dst_channels]
 * drops extra array channels; tensor [height,width,src_channels] -> array region pads RGB with alpha = 1 (`one_bits` = the value
 * 1 in the element type, texture.py:379-380), repeats a single channel (:377-378) and truncates wider data (:383-384). */
int srx_array_to_tensor_ch(void *cuda_array, void *dst_dev, int width, int height, int elem_bytes, int dst_channels, int flip, void *stream);
int srx_tensor_to_array_ch(void *cuda_array, const void *src_dev, int width, int height, int elem_bytes, int src_channels, int flip,
                           int x_offset, int y_offset, unsigned int one_bits, void *stream);
/* CorrespondMap.load (corrmap.py:443-489) without the host round trip: layer <- fp16 atlas [height*width, channels];
 * transpose = 1 reproduces the reference's `get_map(i, order='whc')` upload: texel (x, y) = values[x * width + y]. */
int srx_atlas_to_array(void *cuda_array, const void *values_dev, int height, int width, int channels, int transpose, void *stream);
/* cudaArray of layer `layer` of a mapped GL_TEXTURE_2D_ARRAY resource */
int srx_gl_mapped_layer(srx_gl_resource *res, int layer, void **cuda_array_out);
/* test helpers: plain cudaArray allocation so that the copy kernels can be exercised without a GL context */
int srx_array_alloc(void **cuda_array_out, int width, int height, int channels, int bits_per_channel, int kind /*0 sint,1 uint,2 float*/);
int srx_array_free(void *cuda_array);

/* ------------------------------------------------------------------------------------------------------------------
 * G-buffer frame ingest — replaces RenderManager._save_frame_data (source/engine/managers/renderManager.py:877-948) and the
 * "closer pixel wins" merge of identical-G-buffer draws (renderManager.py:121-133).  Attachments are linear device buffers in
 * the layout of Texture.tensor() (texture.py:166-254): colour / normal+depth / noise RGBA16F, ids RGBA_32I, position / canny
 * RGB32F (renderManager.py:206-367); flip_rows = the buffers are still in GL row order (origin bottom-left).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct srx_gbuffer {
    const void *color;         /* [H,W,4] fp16; alpha = coverage */
    const void *ids;           /* [H,W,4] int32 (spriteID, materialID, map_index, vertexID) */
    const float *pos;          /* [H,W,3] f32 */
    const void *normal_depth;  /* [H,W,4] fp16: normal*0.5+0.5, reversed depth */
    const void *noise;         /* [H,W,4] fp16 */
    const void *canny;         /* [H,W,3] fp16 (the reference's tensor dtype: data_type HALF, renderManager.py:353) or f32 */
    int canny_dtype;           /* SRX_F16 | SRX_F32; canny_maps has the same dtype */
} srx_gbuffer;

typedef struct srx_ingest_args {
    srx_gbuffer src;           /* color is required; the others may be NULL when their outputs are NULL */
    int height, width;         /* multiples of 8 */
    int flip_rows;
    int64_t frame_slot;        /* which frame of the batch tensors below this frame becomes */
    const float *bg_noise;     /* [H,W,4] f32 RenderManager.GlobalBGNoise (:869-875), top-left origin */
    void *color_maps;          /* fp16 [F,H,W,3] (:884)            any output may be NULL = not wanted */
    void *masks;               /* fp16 [F,H,W] = 1 - alpha (:883) */
    int32_t *id_maps;          /* [F,H,W,4] (:895) */
    float *pos_maps;           /* [F,H,W,3] (:902) */
    void *normal_maps;         /* fp16 [F,H,W,3] (:910) */
    void *depth_maps;          /* fp16 [F,H,W,3]: depth repeated three times (:911-912) */
    void *canny_maps;          /* [F,H,W,3] in the attachment's dtype (:942) */
    float *noise_maps;         /* f32 [F,4,H/8,W/8] (:925-937): mask mix, mean of 64 consecutive pixels, AdaIN vs the raw attachment */
    void *workspace;           /* srx_ingest_workspace_bytes(height, width), 256-byte aligned; needed when noise_maps != NULL */
    int64_t workspace_bytes;
} srx_ingest_args;
int64_t srx_ingest_workspace_bytes(int height, int width);
int srx_frame_ingest(const srx_ingest_args *args, void *stream);

typedef struct srx_gbuffer_temp {   /* RenderManager._*_buffer_temp (renderManager.py:219-357), top-left origin; NULL = not kept */
    void *color;               /* fp16 [H,W,4] */
    int32_t *ids;              /* [H,W,4] */
    float *pos;                /* [H,W,3] */
    void *normal;              /* fp16 [H,W,3] */
    void *depth;               /* fp16 [H,W] (required) */
    void *noise;               /* fp16 [H,W,4] */
    void *canny;               /* fp16 [H,W,3] */
} srx_gbuffer_temp;
/* The same attachments as mapped cudaArrays (cudaArray_t from srx_gl_map; NULL = absent): the zero-copy form of the ingest —
 * the kernels read the texture memory through surface objects, nothing is staged (replaces the seven Texture.tensor() read-backs
 * of renderManager.py:882-943: map, Memcpy2D, torch.cuda.synchronize(), flip, clone).  CUDA arrays hold 1, 2 or 4 channels:
 * position / canny (RGB32F in GL, renderManager.py:268,352) are four-channel arrays whose last channel is ignored. */
typedef struct srx_gbuffer_arrays {
    void *color;               /* RGBA16F */
    void *ids;                 /* RGBA32I */
    void *pos;                 /* RGBA32F (xyz used) */
    void *normal_depth;        /* RGBA16F */
    void *noise;               /* RGBA16F */
    void *canny;               /* RGBA32F or RGBA16F (xyz used) */
    int canny_dtype;           /* SRX_F32 | SRX_F16; canny_maps has the same dtype */
} srx_gbuffer_arrays;
/* srx_frame_ingest with `arrays` as the source (args->src is ignored). */
int srx_frame_ingest_arrays(const srx_ingest_args *args, const srx_gbuffer_arrays *arrays, void *stream);
int srx_gbuffer_merge_closer_arrays(const srx_gbuffer_arrays *cur, int height, int width, int flip_rows, const srx_gbuffer_temp *temp,
                                    void *stream);
/* Where cur's reversed depth > temp->depth, every attachment of the pixel replaces the stored one (in place). */
int srx_gbuffer_merge_closer(const srx_gbuffer *cur, int height, int width, int flip_rows, const srx_gbuffer_temp *temp, void *stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Wide-channel feature overlap — the body of OverlapCorresponder.post_atten_inject
 * (source/common_utils/stable_render_utils/corresponder.py:236-295; dead in the reference behind `return origin_values`, :228):
 * group mean of the [B, h*w, c] post-attention features by vertex id through nearest up-sampling to (map_height, map_width),
 * blend with `ratio`, last-writer write-back, nearest down-sampling, AdaIN of the original features to the result's
 * per-(frame, channel) statistics.  (map_height, map_width) = (IDMap.height, IDMap.width) in the reference, which for
 * [F,H,W,4] ids are (W, 4) — corrmap.py:85-93 — the caller decides.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct srx_feature_args {
    const void *ids_dev;        /* [F,H,W,4] */
    int id_dtype;               /* SRX_I32 | SRX_I16 */
    int frames, height, width;
    const int32_t *frame_map_dev; /* DEVICE [F]: batch index of every id frame (IDMap.frame_indices, corresponder.py:254) */
    const void *feat_dev;       /* [B, h*w, c], contiguous */
    void *out_dev;              /* [B, h*w, c], same dtype (may not alias feat_dev) */
    int x_dtype;                /* SRX_F32 | SRX_F16 | SRX_BF16 (computed in fp32, rounded once) */
    int batch, lat_h, lat_w, channels;   /* c <= 1280, c * element size a multiple of 16 bytes */
    int map_height, map_width;
    float ratio;                /* post_attn_inject_ratio (corresponder.py:176, default 0.6) */
    int64_t key_capacity;       /* vertex ids must be < key_capacity */
    void *workspace;            /* srx_feature_overlap_workspace_bytes(), 256-byte aligned */
    int64_t workspace_bytes;
    int reuse_buckets;          /* 0: bucket the ids (one pass + one sort) and apply.  1: `workspace` still holds the buckets of an
                                 * earlier call with the same ids, frame map, sizes (batch, lat_*, map_*) and key capacity — the ids are
                                 * not read; channels, dtype, ratio and the features are free to differ (every attention layer and
                                 * every denoise step of a sampling run sees the same ids).  Only the first
                                 * srx_feature_overlap_bucket_bytes() bytes of the workspace are needed then. */
} srx_feature_args;
int64_t srx_feature_overlap_workspace_bytes(const srx_feature_args *args);
int64_t srx_feature_overlap_bucket_bytes(const srx_feature_args *args);
int srx_feature_overlap(const srx_feature_args *args, void *stream);
/* SRX_ERR_INDEX / SRX_ERR_KEY_RANGE for device-side failures of the bucketing pass behind this workspace (syncs) */
int srx_feature_overlap_check(const srx_feature_args *args, void *stream);
/* feature rows the last call on this workspace gathered (its L2-side work; syncs); -1 on failure */
int64_t srx_feature_overlap_rows(const srx_feature_args *args, void *stream);

/* Cell-similarity overlap — taichi_cells_overlap (source/common_utils/stable_render_utils/corr_utils.py:110-134): every cell
 * becomes the similarity-weighted mean of all cells, similarity = sum over pixel pairs with identical id 4-tuples of the product of
 * their contributions; a cell = pixels / cells consecutive pixels of the flattened frame.  Evaluated in the factorised form
 * (per-key sums), two calls: srx_cells_overlap_keys counts the distinct id tuples (syncs; *n_keys_out on the host), the caller
 * zero-allocates key_sums [n_keys * (channels + 1)] floats, srx_cells_overlap fills new_values (its content is added to, like the
 * reference's placeholder).  ids [batch, pixels, 4] int32, contributions [batch, pixels] f32, values / new_values [batch, cells, c]. */
int64_t srx_cells_overlap_workspace_bytes(int batch, int pixels);
int srx_cells_overlap_keys(const int32_t *ids_dev, const float *contrib_dev, int batch, int pixels, int cells, void *workspace,
                           int64_t workspace_bytes, int64_t *n_keys_out, void *stream);
int srx_cells_overlap(const float *values_dev, float *new_values_dev, int batch, int pixels, int cells, int channels, void *workspace,
                      float *key_sums_dev, int64_t n_keys, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SRX_H_ */
