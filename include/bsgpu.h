/*
 * bsgpu.h -- flat C ABI of libbsgpu.so: B200 (sm_100a) implementation of the two hot
 * paths of JaneliaSciComp/BigStitcher-Spark.
 *
 * Reference call sites each entry point replaces (paths relative to the reference root,
 * J/ = src/main/java/net/preibisch/bigstitcher/spark/):
 *
 *   bs_pcm_*   <-  TransformationTools.computeStitching(...)  J/SparkPairwiseStitching.java:247-255
 *                  (inner numeric cut: PairwiseStitching.getShift -> PhaseCorrelation2.calculatePCM
 *                   + PhaseCorrelation2.getShift, BigStitcher 2.5.0, pom.xml:107);
 *                  parameters: PairwiseStitchingParameters J/SparkPairwiseStitching.java:200-202;
 *                  found == 0  <=>  Java `null` ("No shift found", :274-279).
 *   bs_fuse_*  <-  BlkAffineFusion.initWithIntensityCoefficients(...)  J/SparkAffineFusion.java:602-615
 *                  + BlockAlgoUtils.arrayImg(blockSupplier, interval)   J/SparkAffineFusion.java:620-627
 *                  (multiview-reconstruction 8.0.0 / imglib2-algorithm 0.18.2, pom.xml:101,106);
 *                  dtype converters J/SparkAffineFusion.java:493-517.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; bs_last_error(ctx) gives the text.
 *     No exception crosses the boundary.  There is NO CPU fallback: without a CUDA device
 *     bs_init fails with BS_ERR_CUDA.
 *   - volumes are dense, x-fastest (imglib2 flat order), zero-min; dims are {x, y, z}.
 *   - the caller owns every host buffer; the library owns device memory it allocates.
 *   - a bs_ctx is bound to one device and one compute stream; calls on the same ctx are
 *     serialised by an internal mutex, different ctxs are independent (one per Java worker
 *     thread or one per device; the Spark RDD collapses to a host work queue over ctxs).
 */
#ifndef BSGPU_H
#define BSGPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define BS_OK            0
#define BS_ERR_ARG      -1   /* invalid argument */
#define BS_ERR_CUDA     -2   /* CUDA runtime error / no device */
#define BS_ERR_NOMEM    -3   /* device or host allocation failed */
#define BS_ERR_UNSUPPORTED -4

#define BS_DTYPE_U16 0
#define BS_DTYPE_F32 1
#define BS_DTYPE_U8  2

/* FusionType ordinals (mvrecon FusionGUI.FusionType; CLI help J/SparkAffineFusion.java:124-125) */
#define BS_FUSE_AVG                 0
#define BS_FUSE_AVG_BLEND           1   /* reference default */
#define BS_FUSE_AVG_CONTENT         2
#define BS_FUSE_AVG_BLEND_CONTENT   3
#define BS_FUSE_MAX_INTENSITY       4
#define BS_FUSE_LOWEST_VIEWID_WINS  5
#define BS_FUSE_HIGHEST_VIEWID_WINS 6
#define BS_FUSE_CLOSEST_PIXEL_WINS  7

typedef struct bs_ctx bs_ctx;

/* ---------------------------------------------------------------- lifecycle */
int bs_version(void);
/* device: CUDA ordinal.  stream: an existing cudaStream_t to launch on (e.g. the caller's
 * framework stream), or NULL to let the context create its own non-blocking stream. */
int bs_init(bs_ctx** out, int device, void* stream);
void bs_destroy(bs_ctx* ctx);
const char* bs_last_error(bs_ctx* ctx);   /* ctx may be NULL: last bs_init error of this thread */
int bs_synchronize(bs_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
long long bs_launch_count(bs_ctx* ctx);

/* per-kernel device timing with CUDA events on the context's stream.  Enabling inserts an
 * event pair around every kernel launch; bs_profile_get returns the accumulated milliseconds
 * and launch count for a kernel tag ("fft_x_r2c", "fft_y", "fft_z_xpower", "fft_y_inv",
 * "fft_x_c2r", "peaks", "pearson", "fuse", "content_gauss"). */
int bs_profile_enable(bs_ctx* ctx, int on);
int bs_profile_reset(bs_ctx* ctx);
int bs_profile_get(bs_ctx* ctx, const char* tag, double* ms_total, long long* launches);

/* pinned host memory helpers (the JNI side wraps them in direct ByteBuffers) */
int bs_host_alloc(bs_ctx* ctx, unsigned long long bytes, void** out);
int bs_host_free(bs_ctx* ctx, void* p);

/* ---------------------------------------------------------------- hot path 1: phase correlation */
typedef struct {
    int    peaks_to_check;      /* --peaksToCheck, default 5 (J/SparkPairwiseStitching.java:79-80) */
    int    do_subpixel;         /* !--disableSubpixelResolution (J/SparkPairwiseStitching.java:82-83) */
    int    interpolate_xcorr;   /* PairwiseStitchingParameters.interpolateCrossCorrelation; must be 0 */
    double min_overlap_frac;    /* PairwiseStitchingParameters.minOverlap, default 0.25 */
    int    extension[3];        /* blended-mirror extension in px, upstream fills 10 */
} bs_pcm_params;

typedef struct {
    int       found;            /* 0 <=> Java null */
    long long shift_int[3];     /* integer shift s: img1[p + s] <-> img2[p], {x,y,z} */
    double    shift_sub[3];     /* s + sub-pixel offset (== s when !do_subpixel) */
    double    r;                /* Pearson cross-correlation of the winning candidate */
    long long n_overlap_px;     /* its overlap voxel count */
    long long peak_index[3];    /* PCM index of the winning peak */
    double    pcm_value;        /* PCM value of that peak */
    int       pad[3];           /* padded FFT size used */
    int       n_candidates;     /* Pearson-verified candidates (>= min overlap) */
    long long pearson_px;       /* sum of their overlap voxel counts (byte accounting) */
} bs_pcm_result;

void bs_pcm_default_params(bs_pcm_params* p);

/* one overlap-cropped pair, equal dims (PairwiseStitching.getShift returns null otherwise).
 * on_device != 0: img1/img2 are device pointers on ctx's device (resident data);
 * on_device == 0: host pointers, copied H2D inside the call. */
int bs_pcm_pair(bs_ctx* ctx, const void* img1, const void* img2, const long long dims[3],
                int dtype, const bs_pcm_params* params, int on_device, bs_pcm_result* out);

/* n pairs (all with their own dims[n][3]); host inputs are staged through a double-buffered
 * H2D pipeline on a second stream so copies overlap the previous pair's kernels. */
int bs_pcm_batch(bs_ctx* ctx, int n, const void* const* img1, const void* const* img2,
                 const long long* dims /* n*3 */, int dtype, const bs_pcm_params* params,
                 int on_device, bs_pcm_result* out /* n */);

/* Pairs given as resident volumes plus the raster overlap intervals PairwiseStitching.getShift derives from the
 * two translations (BigStitcher 2.5.0; call site J/SparkPairwiseStitching.java:247-255): a tile is uploaded ONCE
 * (bs_volume_upload[_async]) and takes part in up to 26 pairs; the overlap crops are cut on the device. */
typedef struct {
    unsigned long long vol1, vol2;   /* resident volumes (same dtype) */
    long long min1[3], min2[3];      /* first voxel of the overlap in each volume's own pixel coordinates */
    long long dims[3];               /* overlap size (equal for both; getShift returns null otherwise) */
} bs_pcm_job;
int bs_pcm_volumes_batch(bs_ctx* ctx, int n, const bs_pcm_job* jobs, const bs_pcm_params* params, bs_pcm_result* out);

/* padded FFT length policy of this build (smallest 2^a3^b5^c >= n; even when even != 0) */
int bs_good_fft_size(int n, int even);

/* diagnostic: compute only the PCM of one pair into a host float buffer of pad[0]*pad[1]*pad[2]
 * elements (x-fastest); used by the parity tests to compare spectra-level results. */
int bs_pcm_debug_pcm(bs_ctx* ctx, const void* img1, const void* img2, const long long dims[3],
                     int dtype, const int extension[3], float* out_pcm, int pad_out[3]);

/* ---------------------------------------------------------------- hot path 2: affine fusion */
typedef struct {
    double             src_to_world[12]; /* row-packed 3x4: source pixel -> world, i.e. the adjusted
                                            registration (TransformVirtual.adjustAllTransforms,
                                            J/SparkAffineFusion.java:486-491) times the mipmap transform
                                            (J/util/ViewUtil.java:232-234); inverted by the library */
    unsigned long long vol_handle;       /* resident source volume (bs_volume_upload/_wrap) */
    unsigned long long content_handle;   /* content-weight volume (bs_content_weights) or 0 */
    float              blend_border[3];  /* source px, after FusionTools.adjustBlending */
    float              blend_range[3];
    /* Windowed source (block-wise staging, J/fusion/OverlappingBlocks.java:133-161 / J/util/ViewUtil.java:210-371:
     * only the source cells a block touches are loaded): when full_dims[0] > 0 the resident volume holds the
     * sub-interval [window_min, window_min + volume dims) of a view whose real size is full_dims; src_to_world
     * still maps FULL-view pixel coordinates, the inside test and the blending weights use full_dims, taps are
     * fetched relative to window_min.  The caller guarantees the window covers every tap with non-zero weight
     * (taps outside read zero).  All zeros = the volume is the whole view. */
    long long          full_dims[3];
    long long          window_min[3];
} bs_view;

typedef struct {
    int    fusion_type;     /* BS_FUSE_* */
    int    interpolation;   /* 0 nearest neighbour, 1 n-linear (the reference passes 1) */
    int    out_dtype;       /* BS_DTYPE_F32 / U16 / U8 (J/SparkAffineFusion.java:493-517) */
    int    blend_lut_n;     /* 0: analytic cosine; n>0: n-segment linear-interpolated cosine table */
    double min_intensity;   /* converter range for integer outputs */
    double max_intensity;
    int    out_big_endian;  /* 1: 2- and 4-byte output elements leave the device byte-swapped, i.e. as the big-endian
                             * payload of an N5 block (DefaultBlockWriter), so the host writes the bytes as they come */
    int    reserved;
} bs_fuse_params;

void bs_fuse_default_params(bs_fuse_params* p);

int bs_volume_upload(bs_ctx* ctx, const void* host, const long long dims[3], int dtype,
                     unsigned long long* handle);
/* same, but asynchronous: `host` must be pinned (bs_host_alloc) and stay valid until the copy has run; the copy is
 * queued on the context's copy stream and every later call that uses the handle waits for it on the device, so
 * tile uploads overlap the kernels of earlier work (no host synchronisation).  Device buffers of volumes created
 * this way are recycled through a per-context pool by bs_volume_free. */
int bs_volume_upload_async(bs_ctx* ctx, const void* host, const long long dims[3], int dtype,
                           unsigned long long* handle);
/* register device memory owned by the caller (not freed by bs_volume_free) */
int bs_volume_wrap(bs_ctx* ctx, const void* dev, const long long dims[3], int dtype,
                   unsigned long long* handle);
int bs_volume_free(bs_ctx* ctx, unsigned long long handle);
/* c = G_sigma2 * (I - G_sigma1 * I)^2 on the source volume -> new float32 volume handle */
int bs_content_weights(bs_ctx* ctx, unsigned long long vol_handle, double sigma1, double sigma2,
                       unsigned long long* content_handle);
/* dims {x,y,z} and dtype of a resident volume (either may be NULL) */
int bs_volume_info(bs_ctx* ctx, unsigned long long handle, long long dims[3], int* dtype);
/* copies the whole volume to `host`; capacity_bytes is the size of the caller's buffer and must be at least the
 * volume's byte size (no silent overflow) */
int bs_volume_download(bs_ctx* ctx, unsigned long long handle, void* host, unsigned long long capacity_bytes);
/* next row 8f-3: one 2x half-pixel averaging pyramid step on a resident volume (factors 1 or 2 per
 * axis, output dims = floor(dims / factors), same dtype) -> new handle.  Replaces re-reading level
 * l-1 from the container for every pyramid level (J/SparkAffineFusion.java:703-782). */
int bs_downsample(bs_ctx* ctx, unsigned long long vol_handle, const int factors[3],
                  unsigned long long* out_handle);
/* device address of a resident volume, so that a second context on the same device (another worker
 * thread) can bs_volume_wrap it instead of uploading the tile twice */
int bs_volume_devptr(bs_ctx* ctx, unsigned long long handle, void** dev);

/* fuse one output block: voxel (i,j,k) is at world block_min + (i,j,k)
 * (block_min = gridBlock[0] + bbMin, J/SparkAffineFusion.java:520-534).  views must be in
 * ascending ViewId order.  out: block_size[0]*[1]*[2] elements of out_dtype, x-fastest;
 * out_on_device selects a device or host destination. */
int bs_fuse_block(bs_ctx* ctx, const bs_view* views, int n_views, const long long block_min[3],
                  const long long block_size[3], const bs_fuse_params* params,
                  void* out, int out_on_device);

/* the same for a LIST of blocks in one call (one plan pass, one kernel launch for the whole list; host
 * destinations are staged through two device buffers so the D2H copies of one group of blocks overlap the
 * fusion of the next).  block_min / block_size: n_blocks x 3; outs: n_blocks destination pointers, each a
 * dense x-fastest block like bs_fuse_block's.  This is the work-queue form of the reference's
 * rdd.map(gridBlock -> fuse + save) (J/SparkAffineFusion.java:480-482, 602-670). */
int bs_fuse_blocks(bs_ctx* ctx, const bs_view* views, int n_views, int n_blocks, const long long* block_min,
                   const long long* block_size, const bs_fuse_params* params, void* const* outs, int out_on_device);

/* same, but the fused block stays on the device as a new resident volume (handle): the pyramid
 * levels are then derived with bs_downsample before anything is downloaded (next row 8f-3) */
int bs_fuse_block_to_volume(bs_ctx* ctx, const bs_view* views, int n_views, const long long block_min[3],
                            const long long block_size[3], const bs_fuse_params* params,
                            unsigned long long* out_handle);

/* `affine-fusion --masks [--maskOffset x,y,z]` (J/fusion/GenerateComputeBlockMasks.java:84-176, called at
 * J/SparkAffineFusion.java:564-578): instead of fusing, every block voxel becomes "on" (255 / 65535 / 1.0f for
 * U8 / U16 / F32) when its back-projection into ANY view lies inside [0 - mask_offset, dim - 1 + mask_offset]
 * (source pixels, all three axes).  Only view geometry is used: src_to_world and full_dims (or, when full_dims is 0,
 * the dims of vol_handle).  Same block-list form and output conventions as bs_fuse_blocks. */
int bs_mask_blocks(bs_ctx* ctx, const bs_view* views, int n_views, int n_blocks, const long long* block_min,
                   const long long* block_size, const double mask_offset[3], int out_dtype, int out_big_endian,
                   void* const* outs, int out_on_device);

/* view-sharded mode: accumulate this context's views into partial sums
 * sum_wi / sum_w (device float32, block_size elements each, NOT cleared), to be all-reduced
 * across devices by the caller (NCCL) and finished with bs_fuse_finish. */
int bs_fuse_accumulate(bs_ctx* ctx, const bs_view* views, int n_views, const long long block_min[3],
                       const long long block_size[3], const bs_fuse_params* params,
                       float* sum_wi_dev, float* sum_w_dev);
int bs_fuse_finish(bs_ctx* ctx, const float* sum_wi_dev, const float* sum_w_dev, long long n,
                   const bs_fuse_params* params, void* out, int out_on_device);
/* The exchange itself, inside the library: rank 0 draws a 128-byte NCCL id (bs_comm_unique_id) and hands it to the
 * other ranks by any host channel (the Spark driver / torch.distributed store); every rank joins with bs_comm_init;
 * bs_fuse_allreduce sums both partial buffers of the (overlap) region over all ranks in place -- one grouped NCCL
 * all-reduce over NVLink / NVSwitch, queued on the context's stream between bs_fuse_accumulate and bs_fuse_finish. */
int bs_comm_unique_id(unsigned char id[128]);
int bs_comm_init(bs_ctx* ctx, int n_ranks, int rank, const unsigned char id[128]);
int bs_comm_destroy(bs_ctx* ctx);
int bs_fuse_allreduce(bs_ctx* ctx, float* sum_wi_dev, float* sum_w_dev, long long n);

/* ---------------------------------------------------------------- next row: DoG interest points
 * DoGImgLib2.computeDoG on one block of a resident view (J/SparkInterestPointDetection.java:469-566; the reference
 * passes dog.cuda = null at :490-493 -- this is the device implementation behind that hook).  The caller walks the
 * reference's block grid and expands every block by one voxel inside the image (:397-424). */
typedef struct {
    double sigma;           /* -s, e.g. 1.8 */
    double threshold;       /* -t, e.g. 0.008 */
    double min_intensity;   /* -i0 */
    double max_intensity;   /* -i1 */
    int    find_max;        /* --type MAX / BOTH */
    int    find_min;        /* --type MIN / BOTH */
    int    localization;    /* 0 NONE, 1 QUADRATIC */
    int    pad;
} bs_dog_params;

typedef struct {
    double    loc[3];       /* sub-pixel location {x,y,z} in the view's pixel coordinates */
    double    value;        /* (interpolated) DoG value */
    long long voxel[3];     /* integer location of the extremum */
    int       is_max;
    int       pad;
} bs_dog_point;

void bs_dog_default_params(bs_dog_params* p);
/* detections of the block [interval_min, interval_min + interval_size) sorted by (z, y, x); *n_found may exceed
 * max_points (buffer too small: only max_points were written) */
int bs_dog_detect(bs_ctx* ctx, unsigned long long vol_handle, const long long interval_min[3], const long long interval_size[3],
                  const bs_dog_params* params, bs_dog_point* out, int max_points, int* n_found);

#ifdef __cplusplus
}
#endif
#endif /* BSGPU_H */
