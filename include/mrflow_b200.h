/*
 * mrflow_b200.h -- C ABI of libmrflow_b200.so: the dense plane+parallax hot path of MR-Flow
 * (jswulff/mrflow) as hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI for this path: its boundary is the Python function signature, called
 * in-process with numpy arrays.  The one real FFI in the reference,
 * extern/eff_trws/eff_trws.pyx:4-58 (caller-owned C-contiguous buffers passed as raw pointers,
 * results written into caller-provided arrays), is the model followed here.  Each entry point
 * below names the reference function(s) (file:line under /root/reference) it replaces; the
 * Python host layer in mrflow_b200/ re-exposes them under the reference's own names/signatures.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *  - nothing is allocated inside the library; scratch buffers are caller-provided and their
 *    sizes come from the mrf_*_scratch_bytes() helpers;
 *  - every call is asynchronous on `stream` (a cudaStream_t); no global state, thread-safe per
 *    stream; return value 0 = OK, otherwise a cudaError_t (>0) or an MRF_E_* code (<0); a
 *    message for the calling thread's last failure is available from mrf_last_error();
 *  - images are row-major; `rows`/`y0` describe a row band: the buffers hold `rows` rows whose
 *    first row is global row `y0` of a frame `frame_h` rows tall (y0=0, rows=frame_h for a whole
 *    frame).  Pixel coordinates used in the arithmetic are always the GLOBAL ones, so a band
 *    computes bit-identical values to the whole frame;
 *  - 3x3 matrices are fp64 row-major [9]; epipoles are fp64 [2] = (x, y).
 */
#ifndef MRFLOW_B200_H
#define MRFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mrf_stream_t; /* cudaStream_t */

#define MRF_E_BADARG   (-1)
#define MRF_E_SCRATCH  (-3)

/* sampling modes: the reference samples with cv2.remap(INTER_CUBIC, BORDER_REPLICATE) on
 * 1/32-px quantised coordinates (structure_refinement/interpolation.py:76-80); LINEAR is the
 * same quantisation with INTER_LINEAR (north_star's "bilinear"; rigidity/occlusions.py). */
#define MRF_INTERP_CUBIC   0
#define MRF_INTERP_LINEAR  1

/* robust functions, utils/robust_functions.py:26-70 (argument is x^2), fixed-parameter forms */
#define MRF_RHO_LORENTZIAN    0  /* log(1 + x/sigma^2)            :52 */
#define MRF_RHO_CHARBONNIER   1  /* sqrt(x + eps^2)               :33  (param = eps) */
#define MRF_RHO_GEMAN_MCCLURE 2  /* sigma*x/(x + sigma^2)         :64 */
#define MRF_RHO_QUADRATIC     3  /* x                             :46 */
#define MRF_RHO_LORENTZIAN_AUTO 4 /* s*log(1 + 0.5x/s^2), s given :58  (auto-sigma form) */
#define MRF_RHO_NONE          5  /* the residual Iz itself (for the auto-sigma pre-pass) */

/* cost-volume output layouts */
#define MRF_LAYOUT_LHW_F32 0  /* [L][rows][w] fp32, coalesced along x */
#define MRF_LAYOUT_HWL_I32 1  /* [rows][w][L] int32 = (int)(cost*i32_scale), what
                                 extern/eff_trws/image.cpp:80 consumes */

/* Device-resident whole-frame scalars ("state", fp64 words) written by mrf_cost_volume_front and read
 * by mrf_cost_volume / mrf_structure_to_flow when their dev_state argument is non-NULL. */
#define MRF_STATE_MU      0   /* [2] mu of frame 0 (bwd), 1 (fwd)      compute_structure.py:84 */
#define MRF_STATE_B       2   /* [2] B_b, B_fwd                         compute_structure.py:190-223 */
#define MRF_STATE_MINMAX  4   /* [4] (min, max) of the range-masked structures of frame 0, 1 */
#define MRF_STATE_NANFLAG 8   /* [2] != 0: the range-masked structure of frame 0 / 1 contains a NaN */
#define MRF_STATE_LABELS  16  /* [n_labels] structure_range             build_cost_volume.py:190-192 */
#define MRF_STATE_WORDS   (16 + 256)

int         mrf_version(void);
const char* mrf_last_error(void);

/* ---- a1 + a2: utils/flow_homography.py:4-16 (get_residual_flow) fused with
 *      pipeline/compute_structure.py:517-537 (flow2parallax).
 * u,v fp32 [rows,w]; Hinv = inverse homography; q = epipole.  Any output may be NULL. */
int mrf_flow_to_parallax(const float* u, const float* v, int rows, int w, int y0,
                         const double* Hinv_host, const double* q_host,
                         float* u_res, float* v_res, double* parallax, double* dists,
                         mrf_stream_t stream);

/* ---- a2 on its own: pipeline/compute_structure.py:517-537 (flow2parallax) of a given flow (fp64; an
 * fp32 residual promoted, as numpy promotes fp32 * fp64).  foe_vectors fp64 [rows,w,2]. */
int mrf_flow2parallax(const double* u, const double* v, int rows, int w, int y0, const double* q_host,
                      double* parallax, double* foe_vectors, double* dists, mrf_stream_t stream);

/* ---- mu of pipeline/compute_structure.py:84: sum over the band of ||p - q|| (fp64,
 * deterministic two-stage reduction); out_sum is ONE double on the device.
 * scratch: mrf_reduce_scratch_bytes() bytes. */
size_t mrf_reduce_scratch_bytes(void);
int mrf_epipole_dist_sum(int rows, int w, int y0, const double* q_host,
                         double* out_sum, void* scratch, mrf_stream_t stream);

/* ---- a3: pipeline/compute_structure.py:25-33 (parallax2normedstructure). */
int mrf_parallax_to_structure(const double* parallax, int rows, int w, int y0,
                              const double* q_host, double B, double mu,
                              double* structure, mrf_stream_t stream);

/* ---- a6: pipeline/structure2flow.py:14-38 (structure_to_flow) incl. get_full_flow
 * (utils/flow_homography.py:20-37); if nonrigid != NULL also the paste of
 * combine_flow_simple (:72-73): where nonrigid[p] != 0 the output is flow_init.  With dev_state
 * non-NULL, mu and b are read from the device state of frame state_frame instead. */
int mrf_structure_to_flow(const double* structure, int rows, int w, int y0,
                          const double* q_host, double mu, const double* H_host, double b,
                          const uint8_t* nonrigid, const float* init_u, const float* init_v,
                          float* u, float* v, const double* dev_state, int state_frame,
                          mrf_stream_t stream);

/* ---- a6 tail on its own: utils/flow_homography.py:20-37 (get_full_flow) and :41-48
 * (homography2flow = get_full_flow of a zero residual).  u_res, v_res fp64 (an fp32 residual
 * promoted to fp64, which is what numpy's int64 + fp32 does at :32-33). */
int mrf_full_flow(const double* u_res, const double* v_res, int rows, int w, int y0,
                  const double* H_host, float* u, float* v, mrf_stream_t stream);

/* ---- a7: rigidity/inference.py:28-59 (get_unaries_rigid), dtype behaviour included
 * (fp32 flow magnitude / angle, fp64 elsewhere). */
int mrf_rigidity_unaries(const float* u_res, const float* v_res, int rows, int w, int y0,
                         const double* q_host, double sigma, double prior_rigid,
                         double prior_nonrigid, double* prob, mrf_stream_t stream);

/* ---- a4 (dense part 1): pipeline/compute_structure.py:104-153 -- direction unaries for both
 * frames (a7), the small-motion (0.6) / occlusion (0.5) overrides, the occlusion-aware merge
 * and the CNN-weighted threshold.  Inputs: residual flows of both frames, occlusion masks
 * (uint8, !=0 == occluded), CNN rigidity fp64.  Outputs: p_dir fp64, p_thr uint8, and (optional,
 * device int) thr_count = p_thr.sum() for the TooMuchNonRigid gate (:164). */
int mrf_rigidity_direction(const float* ures0, const float* vres0, const float* ures1, const float* vres1,
                           const uint8_t* occ0, const uint8_t* occ1, const double* p_cnn,
                           int rows, int w, int y0, const double* q0_host, const double* q1_host,
                           double sigma_direction, double weight_cnn,
                           double* p_dir, uint8_t* p_thr, int* thr_count, mrf_stream_t stream);

/* ---- a4 (dense part 2): pipeline/compute_structure.py:172-221 -- valid-pixel mask
 * pv = |par0|>0.1 & |par1|>0.1 & p_thr, and the two candidate lists whose medians give
 * B_fwd (:187-191, structure at B=1 over pv) and B_b (:209-221, template/target over pv).
 * The lists are written compacted (order unspecified) to cand_s1 / cand_ratio (capacity
 * rows*w each); *count (device int) receives their length.  B_fwd enters the ratio list, so
 * call with want=1 for cand_s1 first, take the median, then want=2 with B_fwd for cand_ratio. */
int mrf_structure_candidates(const double* par0, const double* par1, const uint8_t* p_thr,
                             int rows, int w, int y0, const double* q0_host, const double* q1_host,
                             double mu0, double mu1, double B_fwd, int want,
                             uint8_t* pv, double* cand, int* count, mrf_stream_t stream);

/* ---- a4 (dense part 3): pipeline/compute_structure.py:374-394 -- structure-consistency
 * rigidity, merge with direction + CNN, unaries for the MRF.  unaries_f64 is [rows,w,2] fp64 =
 * dstack(1-p, p) (:392); unaries_i32 (optional) is the packing of rigidity/inference.py:113,
 * int32(-unaries*1000), [rows,w,2]. */
int mrf_rigidity_structure(const double* S0, const double* S1, const double* p_dir, const double* p_cnn,
                           const uint8_t* occ0, const uint8_t* occ1, int rows, int w,
                           double mu0, double mu1, double sigma_structure, double weight_cnn,
                           double* unaries_f64, int32_t* unaries_i32, mrf_stream_t stream);

/* ---- a4 without a host round trip: the same pieces with the whole-frame scalars (mu of both frames, B_b,
 * B_fwd) read from the device state (MRF_STATE_*) instead of passed from the host, plus the two one-thread
 * steps that turn the device-resident medians into B_fwd / B_b with the reference's fall-backs
 * (pipeline/compute_structure.py:183-223).  mrflow_b200.planeparallax.structure_pass strings them together:
 * the stage then synchronises with the host once, when its results are downloaded.
 *   mrf_structure_candidates_dev: want = 1 (:187, B = 1) or 2 (:209-219, B_fwd = dev_state[MRF_STATE_B+1]);
 *   mrf_abs_deviation_dev: |x - *c_dev| over the first *n_dev elements (:190, the MAD);
 *   mrf_scales_finalize: mode 1 -> dev_state[MRF_STATE_B+1] = 1.426 * *median_dev, or 1.0 if *n_valid_dev < 100
 *       or scale_structure != 1; mode 2 -> dev_state[MRF_STATE_B] = *median_dev, or -B_fwd if *n_valid_dev < 100;
 *   mrf_parallax_to_structure_dev / mrf_rigidity_structure_dev: :25-33 / :374-394 with B, mu from the state. */
int mrf_structure_candidates_dev(const double* par0, const double* par1, const uint8_t* p_thr, int rows, int w, int y0,
                                 const double* q0_host, const double* q1_host, const double* dev_state, int want,
                                 uint8_t* pv, double* cand, int* count, mrf_stream_t stream);
int mrf_abs_deviation_dev(const double* x, const int* n_dev, size_t capacity, const double* c_dev, double* out,
                          mrf_stream_t stream);
int mrf_scales_finalize(double* dev_state, const int* n_valid_dev, const double* median_dev, int mode,
                        int scale_structure, mrf_stream_t stream);
int mrf_parallax_to_structure_dev(const double* parallax, int rows, int w, int y0, const double* q_host,
                                  const double* dev_state, int frame, double* structure, mrf_stream_t stream);
int mrf_rigidity_structure_dev(const double* S0, const double* S1, const double* p_dir, const double* p_cnn,
                               const uint8_t* occ0, const uint8_t* occ1, int rows, int w, const double* dev_state,
                               double sigma_structure, double weight_cnn, double* unaries_f64, int32_t* unaries_i32,
                               mrf_stream_t stream);

/* ---- exact order statistics for the medians of compute_structure.py:190,221 and the
 * auto-sigma of utils/robust_functions.py:23-29.  keys: n fp64 values (not modified);
 * out3[0] = k-th smallest (0-based), out3[1] = (k+1)-th smallest (so the caller forms numpy's
 * mean-of-two-middles for even n; undefined for k = n-1), out3[2] = number of NaNs (numpy's
 * median is NaN if any; NaNs order last as in numpy's sort).
 * scratch: mrf_select_scratch_bytes(n) bytes, 8-byte aligned. */
size_t mrf_select_scratch_bytes(size_t n);
int mrf_select_kth_f64(const double* keys, size_t n, size_t k, double* out3,
                       void* scratch, mrf_stream_t stream);
/* The same selection in steps, for keys sharded over several GPUs: every
 * rank runs begin(k_global); then 8 x { hist on its shard; all-reduce(SUM) the 256 uint64
 * histogram words at scratch + MRF_SELECT_HIST_OFFSET bytes; pick }; then successor on its
 * shard; all-reduce words 3,4 (SUM: #NaN, #keys <= kth) and word 5 (MIN, unsigned: smallest
 * order-key above kth) of the state at scratch; result. */
#define MRF_SELECT_HIST_OFFSET 64
int mrf_select_begin(size_t k, void* scratch, mrf_stream_t stream);
/* host-driven form of the same steps (prefix / k-th key passed by value, results in caller buffers):
 * hist256 = 256 uint64 counts (zeroed by the call) of byte `pass` (0 = most significant) among the
 * keys whose higher bytes equal `prefix`; out3 = { #keys <= kth, #NaN, min order-key > kth }.  Order
 * keys: the IEEE bits with the sign flipped for positives and all bits flipped for negatives;
 * every NaN = 0xFFFF...F. */
int mrf_select_hist_prefix_f64(const double* keys, size_t n, unsigned long long prefix, int pass,
                               unsigned long long* hist256, mrf_stream_t stream);
int mrf_select_successor_key_f64(const double* keys, size_t n, unsigned long long kth_key,
                                 unsigned long long* out3, mrf_stream_t stream);
int mrf_select_hist_f64(const double* keys, size_t n, void* scratch, mrf_stream_t stream);
int mrf_select_pick(void* scratch, mrf_stream_t stream);
int mrf_select_successor_f64(const double* keys, size_t n, void* scratch, mrf_stream_t stream);
int mrf_select_result(size_t k, void* scratch, double* out3, mrf_stream_t stream);
/* abs deviation |x - c| in place helper for MAD (utils/robust_functions.py:23,
 * compute_structure.py:190) */
int mrf_abs_deviation_f64(const double* x, size_t n, double c, double* out, mrf_stream_t stream);

/* ---- masked min/max for the label range of pipeline/build_cost_volume.py:154-169,190-192:
 * min/max over the band of S with the pixels zeroed that the reference zeroes (rigidity rule
 * + the +-10 px box round the epipole).  zero_mask uint8 (!=0 -> value replaced by 0) may be
 * NULL; box = {x0,x1,y0,y1} half-open in global coords or NULL.  out2 = {min, max} device;
 * NaN propagates as in numpy.  scratch: mrf_minmax_scratch_bytes() bytes. */
size_t mrf_minmax_scratch_bytes(void);
int mrf_masked_minmax_f64(const double* S, const uint8_t* zero_mask, int rows, int w, int y0,
                          const int* box_host, double* out2, void* scratch, mrf_stream_t stream);

/* ---- a11 channel construction, structure_refinement/optimize_structure_single_scale.py:350-371
 * (variational_dataterm_grayscale=1): gray = cvtColor(fp32 RGB)*255 (fp64), np.gradient,
 * 3x3 Gaussian blur.  rgb [h,w,3] interleaved, fp32 or (rgb_is_f64 != 0) fp64 as load_data.py:52
 * produces it; fp64 input is rounded to fp32 first, the reference's I.astype('float32').
 * Outputs (either may be NULL): ch64 fp64 [h,w,3] interleaved = dstack(gray, gy, gx);
 * texels fp32 [h,w,4] = (gray, gy, gx, 0) each rounded to fp32 -- the sampler's input
 * (interpolation.py:76 casts the image to fp32).  One launch per frame: gray, gradients and blur of a tile are
 * formed in shared memory.  gray_scratch: unused since then (kept in the signature; may be NULL). */
int mrf_prepare_channels(const void* rgb, int rgb_is_f64, int h, int w, double* ch64, float* texels,
                         double* gray_scratch, mrf_stream_t stream);

/* ---- a8: structure_refinement/interpolate_through_H.py:7-51 (compute_derivs=False) with
 * interpolation.py:76-80.  img fp32 [img_h,img_w,C] interleaved, C in 1..4; xn,yn fp64 [n];
 * J fp32 [n,C]; mask fp64 [n] (1 inside the frame, else 0).  H may be NULL (= identity
 * without the projective division, i.e. plain interp_lin). */
int mrf_warp_sample(const float* img, int img_h, int img_w, int C, const double* H_host,
                    const double* xn, const double* yn, size_t n, int interp,
                    float* J, double* mask, mrf_stream_t stream);

/* ---- a11 + a12: the per-label data cost (computeDataTerm geometry :210-235, sampling :238,
 * residual :269-286, energy term :302) for ONE neighbour frame and ALL labels, i.e. the job
 * of the missing extension at pipeline/build_cost_volume.py:209-219. */
typedef struct {
  /* reference frame band */
  int rows, w, y0, frame_h;
  /* neighbour-frame texel buffer [nb_rows][w][4] fp32 holding global rows nb_y0.. */
  int nb_rows, nb_y0;
  double H[9];          /* homography reference -> neighbour */
  double q[2];          /* epipole */
  double mu, B;
  double label_scale;   /* mu0/mu1 for the backward frame, 1 for the forward one (:435-436) */
  int    n_labels;      /* <= 256 */
  int    interp;        /* MRF_INTERP_* */
  int    rho;           /* MRF_RHO_* */
  double rho_param;     /* sigma or eps */
  double cw[3];         /* channel weights (0.01,1,1) :371 */
  int    layout;        /* MRF_LAYOUT_* */
  double i32_scale;     /* for MRF_LAYOUT_HWL_I32 */
  int    variant;       /* MRF_VARIANT_* */
  int    state_frame;   /* with dev_state: which frame's mu / B to read (0 = bwd, 1 = fwd) */
  long long cost_plane_stride; /* LHW layout: elements between consecutive label planes of `cost`;
                                  0 = rows*w.  With frame_h*w and cost = volume + y0*w a band writes
                                  its rows straight into a whole-frame [L][frame_h][w] volume -- which
                                  may live on another GPU (peer-mapped memory, mrf_peer_*): the gather
                                  to the TRW-S rank is then done by the kernel's own stores over NVLink. */
} mrf_costvol_params;

/* kernel variants: STAGED copies the tile's footprint window of the neighbour frame into shared
 * memory with TMA bulk copies and gathers from there (default); GATHER reads every tap through
 * L1/L2 from global memory.  Results are bit-identical. */
#define MRF_VARIANT_STAGED 0
#define MRF_VARIANT_GATHER 1
#define MRF_VARIANT_OCC2   16 /* OR-able tuning flag: compile for 2 resident blocks / SM (128
                                 registers per thread) instead of 3 */
#define MRF_VARIANT_OCC4   32 /* OR-able: 4 resident blocks / SM (64 registers, 40 KB staged window) */
#define MRF_VARIANT_GROUPED 128 /* OR-able, staged only: the label loop handles two labels per iteration, each phase (fp64
                                   geometry, taps, fp64 residual / rho) one branch-free straight line over both.  Same
                                   arithmetic.  Default for bilinear sampling (8 % faster); opt-in for bicubic (measured
                                   10 % slower: the kernel is bound by shared-memory wavefronts there) */
#define MRF_VARIANT_SINGLE  256 /* OR-able: bilinear sampling with the one-label-per-iteration loop (for comparison) */

/* ref_ch: fp64 [rows,w,3] channels of the reference frame band; nb_texels as described;
 * labels: fp64 [n_labels] structure_range (device); nonrigid/occluded: uint8 [rows,w], !=0
 * zeroes the residual (:285-286), may be NULL.  cost: layout-dependent, may be NULL (then only
 * min/argmin are produced); min_cost fp32 [rows,w] / argmin int32 [rows,w] (first minimum,
 * np.argmin semantics) may be NULL.  halo_miss (device int, may be NULL; zero it first) counts
 * in-frame samples whose taps fall outside the neighbour band nb_y0..nb_y0+nb_rows-1 -- a
 * non-zero value means the caller's halo was too small and the band must be recomputed. */
int mrf_cost_volume(const mrf_costvol_params* p_host, const double* ref_ch, const float* nb_texels,
                    const double* labels, const uint8_t* nonrigid, const uint8_t* occluded,
                    void* cost, float* min_cost, int32_t* argmin, int* halo_miss,
                    const double* dev_state, mrf_stream_t stream);
/* Both neighbour frames of a triplet in ONE launch (grid.z = frame): same arithmetic, but the tail of one
 * frame's blocks overlaps the head of the other's, and there is one launch less.  p0 / p1: the per-frame
 * parameters (H, q, mu, B, label_scale, state_frame, nb band); shape, sampler, rho, layout, variant and
 * n_labels must agree.  ref_ch, labels, nonrigid, halo_miss and dev_state are shared. */
int mrf_cost_volume_pair(const mrf_costvol_params* p0_host, const mrf_costvol_params* p1_host, const double* ref_ch,
                         const float* nb_texels0, const float* nb_texels1, const double* labels,
                         const uint8_t* nonrigid, const uint8_t* occluded0, const uint8_t* occluded1,
                         void* cost0, void* cost1, float* min_cost0, float* min_cost1, int32_t* argmin0,
                         int32_t* argmin1, int* halo_miss, const double* dev_state, mrf_stream_t stream);

/* ---- auto-sigma mode of the cost volume (utils/robust_functions.py:55-58, the reference's default
 * generate_lorentzian()): sigma = max(1e-6, 1.4826 * median |Iz|) over the WHOLE residual plane of a label, so
 * each label plane is built in three steps -- mrf_cost_residual_plane (fp64 residual Iz of
 * optimize_structure_single_scale.py:210-286 and |Iz|; `cost` semantics of mrf_cost_volume apply to the other
 * arguments), mrf_select_median_dev over |Iz|, mrf_cost_auto_rho_plane (sigma * log(1 + 0.5 Iz^2 / sigma^2) as
 * fp32, running min / first argmin over the labels, label = 0, 1, ... in order).  Plain kernels: the reference
 * arithmetic written naively. */
int mrf_cost_residual_plane(const mrf_costvol_params* p_host, const double* ref_ch, const float* nb_texels,
                            const double* labels, const uint8_t* nonrigid, const uint8_t* occluded, int label,
                            double* residual, double* abs_residual, int* halo_miss, const double* dev_state,
                            mrf_stream_t stream);
int mrf_cost_auto_rho_plane(const double* residual, size_t n, const double* median_abs_dev, int label,
                            float* cost_plane, float* min_cost, int32_t* argmin, mrf_stream_t stream);

/* ---- the pre-pass of build_cost_volume with no host round trip (pipeline/build_cost_volume.py:78-192
 * with the live conventions of compute_structure.py): residual flow -> parallax (both frames), mu,
 * valid mask and B_b = median(template/target) with B_fwd fixed, normed structures, masked min/max,
 * label set.  All whole-frame scalars land in dev_state (MRF_STATE_*), so the label-major cost
 * volume can follow on the same stream -- or the whole step be captured in a CUDA graph.  Whole
 * frame on one GPU only (row bands use the step-wise calls with all-reduces in between).
 * rigid_mask: uint8 (the reference's rigidity > 0); zero_mask: pixels zeroed before the range is
 * taken (:159) or NULL; box: the +-10 px square round each epipole {x0,x1,y0,y1} (:154-165).
 * cand: rows*w doubles; count: device int; scratch: mrf_front_scratch_bytes() bytes. */
typedef struct {
  int rows, w, y0, frame_h;
  double Hinv0[9], Hinv1[9];   /* inverse homographies, utils/flow_homography.py:12 */
  double q0[2], q1[2];         /* (refined) epipoles */
  double B_fwd;                /* 1.0 in build_cost_volume.py:131 */
  int have_box[2];
  int box[2][4];
  int n_labels;
  int fused;                   /* 0 (default): nine short launches; 1: the same pre-pass as ONE persistent
                                  cooperative kernel with grid barriers -- identical results.  Measured at Sintel
                                  size (profiles/r2g_*): 92 us alone against ~105 us for the launches, but a
                                  cooperative grid that spins at barriers takes the SMs from the cost-volume
                                  kernels of other triplets in flight (0.68 vs 0.63 ms per triplet), hence opt-in */
} mrf_front_params;
size_t mrf_front_scratch_bytes(void);
int mrf_cost_volume_front(const mrf_front_params* p_host, const float* u0, const float* v0,
                          const float* u1, const float* v1, const uint8_t* rigid_mask,
                          const uint8_t* zero_mask, double* par0, double* par1, double* S0, double* S1,
                          double* cand, int* count, double* dev_state, void* scratch, mrf_stream_t stream);
/* The same pre-pass in pieces for ROW BANDS (p->y0, p->rows = this rank's band).  Between the pieces the
 * caller runs SEVEN small collectives on the same stream (NCCL) -- no host round trip, and none for mu:
 *   mrf_front_mu_whole     -> dev_state[MRF_STATE_MU..]: the sum of ||p - q|| over the WHOLE frame depends
 *                             only on (q, frame_h, w), so every rank evaluates it itself, in the grid and
 *                             summation tree of mrf_cost_volume_front: the bits of the one-GPU run;
 *   mrf_front_parallax_band-> par0, par1 of the band;
 *   mrf_front_candidates_band -> cand / count (local) and the first-digit histogram of the exact median
 *                             (select13: five passes of 12+13+13+13+13 bits);
 *   5 x { all-reduce SUM of the MRF_SELECT13_BINS uint32 words of pass p at select_scratch +
 *         MRF_SELECT13_HIST_OFFSET + p * 4 * MRF_SELECT13_BINS;  mrf_select13_hist_dev(pass p+1) -- after the
 *         fifth all-reduce mrf_select13_successor_dev instead };
 *   all-gather the 3 uint64 words at select_scratch + MRF_SELECT13_SUCC_OFFSET of every rank;
 *   mrf_front_structures_band(gathered) -> dev_state[MRF_STATE_B] = the median, S0, S1, and this band's
 *                             range min / max / NaN flags in dev_state[MRF_STATE_MINMAX .. MRF_STATE_NANFLAG+1];
 *   all-gather those 6 doubles of every rank;  mrf_label_range_dev(gathered). */
#define MRF_SELECT13_BINS        8192
#define MRF_SELECT13_HIST_OFFSET 256
#define MRF_SELECT13_SUCC_OFFSET 24
/* The same exchanges WITHOUT collective launches (csrc/peersync.cuh): with a non-NULL `peer` the kernels add
 * their histogram bins straight into every rank's histogram over NVLink (remote reductions on peer-mapped
 * memory), their last block sends the few gathered words to every rank, publishes the phase with a flag store,
 * and the next kernel polls its own flags -- one one-way NVLink latency per exchange.  Every rank owns a context
 * of mrf_peer_sync_bytes() bytes from mrf_peer_alloc, zeroed once, opened by all peers (ctx[r] = this process's
 * pointer to rank r's context, its own at [rank]).  Call sequence per step, all ranks alike:
 *   mrf_peer_sync_begin; mrf_front_mu_whole / parallax_band; candidates_band(peer); mrf_peer_push_hist(0);
 *   4 x { mrf_select13_hist_dev(pass p, peer); mrf_peer_push_hist(p) }; mrf_select13_successor_dev(peer);
 *   mrf_front_structures_band(gathered NULL, peer); mrf_label_range_dev(gathered NULL, peer).  A rank that never shows up sets the context's error word (second
 * 8-byte word) after 20 s of wall clock; check it before trusting results. */
typedef struct {
  void* ctx[8];
  int world, rank;
} mrf_peer_sync;
size_t mrf_peer_sync_bytes(void);
int mrf_peer_sync_begin(const mrf_peer_sync* peer_host, mrf_stream_t stream);
/* after the kernel that produced this rank's local histogram of `pass` (candidates_band: 0; hist_dev: 1..4):
 * add it into every rank's histogram over NVLink and publish the phase (32 blocks) */
int mrf_peer_push_hist(const mrf_peer_sync* peer_host, const void* select_scratch, int pass, mrf_stream_t stream);
int mrf_front_mu_whole(const mrf_front_params* p_host, double* dev_state, void* scratch, mrf_stream_t stream);
int mrf_front_parallax_band(const mrf_front_params* p_host, const float* u0, const float* v0,
                            const float* u1, const float* v1, double* par0, double* par1,
                            void* scratch, mrf_stream_t stream);
int mrf_front_candidates_band(const mrf_front_params* p_host, const double* par0, const double* par1,
                              const uint8_t* rigid_mask, double* dev_state, double* cand, int* count,
                              void* select_scratch, const mrf_peer_sync* peer_host, mrf_stream_t stream);
int mrf_front_structures_band(const mrf_front_params* p_host, const double* par0, const double* par1,
                              const uint8_t* zero_mask, double* S0, double* S1, double* dev_state,
                              const void* select_scratch, const unsigned long long* gathered_successor, int world,
                              void* scratch, const mrf_peer_sync* peer_host, mrf_stream_t stream);
/* Exact median in five radix passes with the counts read on the device (csrc/select13.cuh): zero the
 * scratch (mrf_select_scratch_bytes), histogram pass 0..4 (each resolves the previous digit from the
 * finished -- for row bands all-reduced -- histogram), successor pass, median (gathered = the successor
 * words of all ranks [world][3], or NULL on one GPU). */
int mrf_select13_begin(void* scratch, mrf_stream_t stream);
int mrf_select13_hist_dev(const double* keys, const int* n_local_dev, size_t capacity, int pass, void* scratch,
                          const mrf_peer_sync* peer_host, mrf_stream_t stream);
int mrf_select13_successor_dev(const double* keys, const int* n_local_dev, size_t capacity, void* scratch,
                               const mrf_peer_sync* peer_host, mrf_stream_t stream);
int mrf_select13_median(const void* scratch, const unsigned long long* gathered, int world, double* out_median,
                        mrf_stream_t stream);
/* numpy.median of the first *n_dev keys (n read on the device); np.linspace label set from the
 * state's min/max.  Building blocks of the front, exported for tests. */
int mrf_select_median_dev(const double* keys, const int* n_dev, size_t capacity, double* out_median,
                          void* scratch, mrf_stream_t stream);
/* gathered: the six words MRF_STATE_MINMAX .. MRF_STATE_NANFLAG+1 of every band [world][6], or NULL */
int mrf_label_range_dev(double* dev_state, int n_labels, const double* gathered, int world,
                        const mrf_peer_sync* peer_host, mrf_stream_t stream);

/* ---- utils/robust_functions.py:26-70 on a device array: out = rho(x) (deriv = 0) or d rho/dx
 * (deriv = 1), x = squared residual.  kind = MRF_RHO_*; p1 = sigma / eps; p2 = Charbonnier
 * exponent a (:26; 0.5 selects the sqrt form) or, for the Geman-McClure derivative, sigma**3 as
 * the caller's Python evaluates it (:65).  MRF_RHO_NONE writes sqrt(x), the sample of the MAD
 * auto-sigma (:23-29). */
int mrf_robust_apply_f64(const double* x, size_t n, int kind, int deriv, double p1, double p2,
                         double* out, mrf_stream_t stream);

/* ---- pipeline/optimize_variational_refinement.py:20-42 (the stage between compute_structure and
 * combine_flow when override_optimization=1): 3x3 median of both occlusion maps (uint8 0/1), then
 * structure = fwd_only*S0' + bwd_only*S1 + none*(S0'+S1)/2 with S0' = S0*mu1/mu0.  Whole frame.
 * occ_*_blur (optional, uint8) receive the median-filtered maps the variational solver is given (:83-84). */
int mrf_merge_structures(const double* S0, const double* S1, const uint8_t* occ_bwd, const uint8_t* occ_fwd,
                         int h, int w, double mu0, double mu1, int occlusion_reasoning, double* out,
                         uint8_t* occ_bwd_blur, uint8_t* occ_fwd_blur, mrf_stream_t stream);

/* ---- pipeline/compute_structure.py:44: cv2.erode(uint8, ones((3,3))) (default border). */
int mrf_erode3x3_u8(const uint8_t* src, int h, int w, uint8_t* dst, mrf_stream_t stream);

/* ---- the stages either side of the path.
 * rigidity/occlusions.py:10-74 (computeOcclusionsFromConsistency): flow[0] = (uf0, vf0) backward,
 * flow[1] = (uf1, vf1) forward, backflow likewise; all fp32 [h,w].  Masks uint8 [h,w]; occ_both may
 * be NULL.  The error is compared as fp32 against (float)threshold, as numpy does. */
int mrf_flow_consistency(const float* uf0, const float* vf0, const float* ub0, const float* vb0,
                         const float* uf1, const float* vf1, const float* ub1, const float* vb1,
                         int h, int w, double threshold, uint8_t* occ_backward, uint8_t* occ_forward,
                         uint8_t* occ_both, mrf_stream_t stream);
/* utils/flow_io.py:33-48 (flow_read), the payload half: file_dev = the WHOLE .flo file as it lies on disk, copied to
 * device memory (4-byte aligned): 12-byte header (fp32 tag 202021.25, int32 width, int32 height -- validate it on the
 * host, as mrflow_b200.utils.flow_io.flow_read does), then height*width interleaved (u, v) fp32 pairs.  u, v: fp32
 * [h,w] planes, bit copies of the payload. */
int mrf_flo_decode(const void* file_dev, size_t file_bytes, int h, int w, float* u, float* v, mrf_stream_t stream);
/* rigidity/inference.py:140-167 (get_image_weights): I fp64 [h,w,c]; four fp32 [h,w] planes in the
 * reference's return order (horizontal, vertical, ne, se), ready for Eff_TRWS.solve
 * (extern/eff_trws/eff_trws.pyx:25-53).  scratch: mrf_image_weights_scratch_bytes() bytes. */
size_t mrf_image_weights_scratch_bytes(void);
int mrf_image_weights(const double* I, int h, int w, int c, float* w_horizontal, float* w_vertical,
                      float* w_ne, float* w_se, void* scratch, mrf_stream_t stream);

/* ---- the linearised data term of the variational refinement.
 * a9, structure_refinement/interpolate_through_H.py:53-97 (compute_derivs_through_H): I fp64 [h,w,C];
 * M4 = the four matrices Hinv*T*H of :68 (x+1, x-1, y+1, y-1), Minv4 = their inverses as cv::invert
 * forms them (closed form), bw0 = OpenCV's warp block width for this frame size; dx, dy fp64 [h,w,C]. */
int mrf_derivs_through_H(const double* I, int h, int w, int C, const double* M4_host,
                         const double* Minv4_host, int bw0, double* dx, double* dy, mrf_stream_t stream);
/* fp64 [h,w,3] -> packed fp32 texels [h,w,4] (interpolation.py:76 casts what it samples to fp32) */
int mrf_pack_texels(const double* src, int h, int w, float* texels, mrf_stream_t stream);
/* optimize_structure_single_scale.py:203-291 (computeDataTerm, first half): per-pixel structure A ->
 * warp -> I2w, dI2w_dx, dI2w_dy (bicubic, through H), averaged with the reference frame's derivatives,
 * channel-weighted; outputs the residual Iz, the directional derivative dI2w_dphi and dr/da. */
int mrf_data_term_residual(const double* A, const double* ref_ch, const float* nb_texels,
                           const float* nb_dx_texels, const float* nb_dy_texels, const double* ref_dx,
                           const double* ref_dy, const uint8_t* nonrigid, const uint8_t* occluded, int h,
                           int w, const double* H_host, const double* q_host, double mu, double B,
                           const double* cw_host, double* Iz, double* dphi, double* dr_da,
                           mrf_stream_t stream);
/* :294-302: psi = rho'(Iz^2), diag = psi*dphi^2*dr_da^2, b = psi*Iz*dphi*dr_da (before the median
 * filters) and the energy sum(rho(Iz^2)) (one double on the device).  rho in MRF_RHO_LORENTZIAN ..
 * MRF_RHO_LORENTZIAN_AUTO; dev_median (optional, device): the median of |Iz| -- the kernel then forms the
 * auto-sigma max(1e-6, 1.4826*median) itself and rho_param is ignored, so the host never has to see it;
 * scratch: mrf_data_term_scratch_bytes() bytes. */
size_t mrf_data_term_scratch_bytes(void);
int mrf_data_term_system(const double* Iz, const double* dphi, const double* dr_da, size_t n, int rho,
                         double rho_param, const double* dev_median, double* diag, double* b, double* energy,
                         void* scratch, mrf_stream_t stream);
/* scipy.ndimage.median_filter(size=5) with its default 'reflect' boundary (:299-300) */
int mrf_median5x5_f64(const double* src, int h, int w, double* dst, mrf_stream_t stream);

/* ---- the variational structure refinement around the data term:
 * structure_refinement/optimize_structure_single_scale.py:310-652 without its scipy.sparse matrices.
 * utils/edge_weights.py:38-55 (computeWeightsLocallyNormalized) on the fp32 gray image + the masking of
 * :388-392 (rigidity_eroded may be NULL: no masking); wx, wy fp32 [h,w]; tmp_a/tmp_b fp32 [h,w] and
 * tmp_ra/tmp_rb fp64 [h,w] are caller-provided work planes. */
int mrf_gray32(const void* rgb, int rgb_is_f64, size_t n, float* gray32, mrf_stream_t stream);
int mrf_edge_weights_local(const float* gray32, const uint8_t* rigidity_eroded, int h, int w, int norm_radius,
                           double min_weight, float* wx, float* wy, float* tmp_a, float* tmp_b,
                           double* tmp_ra, double* tmp_rb, mrf_stream_t stream);
size_t mrf_var_scratch_bytes(void);
/* :505-511, :541-547: the robust arguments of the first / second order smoothness terms (differences
 * with replicated borders, structure_refinement/construct_matrices.py:14-38) */
int mrf_var_smooth_args(const double* A, const float* wx, const float* wy, int h, int w, double afac_sq,
                        double* arg1, double* arg2, double* root1, double* root2, mrf_stream_t stream);
/* Lorentzian weights for given sigmas (:511-512, :547-548) and the two energies (:560-561, two doubles
 * on the device); dev_medians2 (optional, device): medians of root1 / root2, the auto-sigmas are then formed
 * in the kernel */
int mrf_var_smooth_weights(const double* arg1, const double* arg2, const float* wx, const float* wy, size_t n,
                           double afac_sq, double sigma1, double sigma2, const double* dev_medians2,
                           double* p1x, double* p1y, double* p2x, double* p2y, double* energies2, void* scratch,
                           mrf_stream_t stream);
/* :448-461 forward data term with the backward one where the forward frame is occluded, and the
 * consistency residual / argument of :476-490 */
int mrf_var_data_merge(const double* diag_f, const double* b_f, const double* diag_b, const double* b_b,
                       const uint8_t* occ_f, const uint8_t* occ_b, const double* A, const double* A_bwd_ref,
                       const double* A_fwd_ref, size_t n, double scale_bwd, double scale_ref, double afac_sq,
                       double* D, double* bd, double* Iz_c, double* arg_c, double* root_c, mrf_stream_t stream);
/* Charbonnier weight (auto eps from the caller's median) and energy of the consistency term (:491-497) */
int mrf_var_consistency_weights(const double* arg_c, size_t n, double afac_sq, double eps,
                                const double* dev_median, double* pc, double* energy, void* scratch,
                                mrf_stream_t stream);
/* y = Z (D + l1 G'P1G + l2 G2'P2G2 + lc Pc) Z x + eps x  (:580-595), matrix-free; rigid / D / pc may be
 * NULL (identity / absent) */
int mrf_var_apply(const double* x, const uint8_t* rigid, const double* D, const double* p1x, const double* p1y,
                  const double* p2x, const double* p2y, const double* pc, double l1, double l2, double lc,
                  double eps, int h, int w, double* y, mrf_stream_t stream);
/* b = Z (-bd - smooth_A - lc * pc * Iz_c)  (:581, :588) */
int mrf_var_rhs(const double* bd, const double* smooth_A, const double* pc, const double* Iz_c,
                const uint8_t* rigid, size_t n, double lc, double* b, mrf_stream_t stream);
int mrf_vec_scale(const double* y, size_t n, double s, double* out, mrf_stream_t stream);
/* scipy.sparse.linalg.minres (the solver behind :86/:593; x0 = 0, no preconditioner) kept on the device:
 * mrf_minres_iterate is ONE persistent cooperative kernel that runs iterations itn_done+1 ..
 * itn_done+n_iters (or until the stopping tests fire) with two grid barriers per iteration; the scalar
 * recurrences and stopping tests are evaluated in the kernel and kept in ``state``
 * (MRF_MINRES_STATE_WORDS doubles), which makes the solve resumable.  r3 = three vectors [3][n] (r1, r2 = y
 * and the next y, rotating), w2 and v2 = two vectors [2][n] each, x one; scratch:
 * mrf_minres_scratch_bytes(h, w) bytes.  The operator is that of mrf_var_apply. */
#define MRF_MINRES_STATE_WORDS 32
#define MRF_MINRES_ITN    18
#define MRF_MINRES_ISTOP  19   /* scipy's istop (1..6, -1); 100 = b was zero; 7 = non-finite system (stopped at once) */
#define MRF_MINRES_ANORM  20
#define MRF_MINRES_ACOND  21
#define MRF_MINRES_RNORM  22
#define MRF_MINRES_YNORM  23
#define MRF_MINRES_NS_A    29  /* block 0's own time in phase A / phases B+C / barrier waits, nanoseconds, */
#define MRF_MINRES_NS_BC   30  /* accumulated over the launches of a solve (for profiles/, not for control) */
#define MRF_MINRES_NS_SYNC 31
size_t mrf_minres_scratch_bytes(int h, int w);
int mrf_minres_init(const double* b, size_t n, double* r3, double* w2, double* x, double* state, void* scratch,
                    mrf_stream_t stream);
int mrf_minres_iterate(const uint8_t* rigid, const double* D, const double* p1x, const double* p1y,
                       const double* p2x, const double* p2y, const double* pc, double l1, double l2, double lc,
                       double eps, int h, int w, double* r3, double* w2, double* v2, double* x, double* state,
                       void* scratch, int itn_done, int n_iters, double rtol, long long maxiter,
                       mrf_stream_t stream);
/* :600-612 NaN/inf scrub, clip_dA (:157-183), clip to +-1; writes A + dA */
int mrf_var_clip_update(const double* da, const double* A, int h, int w, const double* q_host, double mu,
                        double B, double threshold, double* A_plus_da, mrf_stream_t stream);
/* scipy.ndimage.median_filter(size=7), 'reflect' (:636) and A <- A + (median - A) (:650) */
int mrf_median7x7_f64(const double* src, int h, int w, double* dst, mrf_stream_t stream);
int mrf_var_advance(double* A, const double* median_of_A_plus_da, size_t n, mrf_stream_t stream);

/* ---- peer memory for the row-band gather: a buffer allocated on the
 * TRW-S rank's GPU is exported as a 64-byte CUDA IPC handle, opened by the other ranks (one process
 * per GPU), and written by their cost-volume kernels directly over NVLink.  These are the only
 * calls that allocate; they are not stream-ordered. */
int mrf_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle64_out);
int mrf_peer_open(const unsigned char* handle64, void** ptr_out);
int mrf_peer_close(void* ptr);
int mrf_peer_free(void* ptr);

/* ---- device -> host of a row chunk of a label-major volume: rows [y0, y0+rows) of every plane of a
 * device [n_planes][frame_h][w] array into the same rows of a (pinned) host array of the same shape, as one
 * pitched DMA on `stream` -- the download of chunk k overlaps the kernel of chunk k+1
 * (the 91 MB per neighbour frame that pipeline/build_cost_volume.py:226-231 returns is the long pole of a
 * host-API call). */
int mrf_download_rows(void* dst_host, const void* src_dev, size_t elem_bytes, int n_planes, int frame_h,
                      int w, int y0, int rows, mrf_stream_t stream);

/* ---- small all-reduces over peer memory: integer / min reductions of up to 256
 * words (a building block; the resident row-band pre-pass itself uses NCCL for its 32 KB histograms) as ONE
 * kernel per rank with a low-latency protocol: every 64-bit value travels as two 8-byte words
 * (tag<<32 | half) stored straight into every rank's mailbox over NVLink -- an 8-byte store is one
 * transaction, so a word carrying the call's tag is the data: no fence, no flag; the receiver polls its
 * own mailbox and reduces in rank order.  A peer that has not arrived after 20 s of wall clock
 * (%globaltimer) turns the WHOLE result into all-ones words and sets the mailbox's status word (its last 8
 * words): callers must check it (mrflow_b200.sharding.PeerMailbox.check) before using results.
 * A mailbox is mrf_mailbox_bytes(world) bytes from mrf_peer_alloc, zeroed once, opened by every peer
 * (mrf_peer_open); mailboxes_host[world] are this process's pointers to all of them (its own at [rank]).
 * call_index must count the collectives 0, 1, 2, ... identically on every rank.  inout: nwords <= 256
 * 64-bit words on the device.  Only order-independent reductions, so results equal NCCL's bit for bit. */
#define MRF_PEER_SUM_U64 0
#define MRF_PEER_MIN_U64 1
#define MRF_PEER_MIN_F64 2
size_t mrf_mailbox_bytes(int world);
int mrf_mailbox_allreduce(void* local_mailbox, void* const* mailboxes_host, int world, int rank,
                          long long call_index, int op, void* inout_words, int nwords, mrf_stream_t stream);

/* number of kernel launches the library has issued in this process (for bench.py's
 * gpu_launches); reset with mrf_reset_launch_count(). */
long long mrf_launch_count(void);
void      mrf_reset_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MRFLOW_B200_H */
