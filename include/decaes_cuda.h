/*
 * libdecaes_cuda — C ABI for the B200-native voxelwise T2-distribution pipeline.
 *
 * This header is the drop-in boundary for ONE hot path of DECAES.jl: the per-voxel
 * worker loops of T2mapSEcorr! and T2partSEcorr.  The reference has no FFI for this
 * path (it is pure Julia); each entry point below states the reference code whose
 * body it replaces.  All citations are relative to the reference checkout.
 *
 *   decaes_t2map        <- worker loop of T2mapSEcorr!          src/T2mapSEcorr.jl:177-193
 *                          (voxelwise_T2_distribution!           src/T2mapSEcorr.jl:201-238,
 *                           save_results!                        src/T2mapSEcorr.jl:512-591)
 *                          optionally fused with the T2part epilogue
 *                          (voxelwise_T2_parts!                  src/T2partSEcorr.jl:95-138)
 *   decaes_t2part       <- worker loop of T2partSEcorr           src/T2partSEcorr.jl:58-68
 *   decaes_setup_tables <- T2Maps(opts) table fields             src/T2mapSEcorr.jl:24-33
 *                          (echotimes, t2times, refangleset, decaybasisset)
 *
 * Conventions
 *   - Plain C types only; every array is Float64, column-major exactly as Julia lays
 *     it out: image (nx,ny,nz,nTE) => voxel v, echo e at image[v + e*Nvox], Nvox=nx*ny*nz;
 *     dist (nx,ny,nz,nT2) => dist[v + j*Nvox]; maps (nx,ny,nz) => map[v].
 *   - Host entry points (decaes_t2map / decaes_t2part) take HOST pointers owned by the
 *     caller, are blocking and shard voxel slabs over `ngpus` devices with no collective.
 *     Voxels with image[v,0] <= Threshold are written as NaN — what the reference's
 *     NaN pre-fill (src/T2mapSEcorr.jl:36-52, tfill(NaN)) leaves for skipped voxels;
 *     `alpha` keeps the caller's value there when alpha_provided = 1.
 *   - *_device entry points take DEVICE pointers on the current CUDA device and enqueue
 *     on the given stream (cudaStream_t passed as void*); they do not synchronise.
 *     The library keeps ONE set of per-device tables, scratch and kernel parameters, so
 *     launches on the same device are serialised: a per-device mutex guards the host side and
 *     every new launch waits (cudaStreamWaitEvent) for the previous pipeline kernel of that
 *     device, whatever stream it was enqueued on.  Different devices run concurrently.
 *   - Return value: 0 on success, negative decaes_status on failure; the message is in
 *     decaes_last_error() (thread-local).  Nothing throws across the ABI.
 *   - There is no CPU fallback: without a usable CUDA device every compute entry point
 *     fails with DECAES_ECUDA.
 */
#ifndef DECAES_CUDA_H
#define DECAES_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DECAES_ABI_VERSION 2  /* v2: decaes_run_stats grew the counted-voxel fields (appended; v1 prefix unchanged) */

typedef enum {
  DECAES_OK = 0,
  DECAES_EINVAL = -1,       /* option fails a T2mapOptions/T2partOptions assertion (src/types.jl:28-84,148-168) */
  DECAES_ECUDA = -2,        /* CUDA runtime error / no device */
  DECAES_EUNSUPPORTED = -3, /* size outside the accelerated path (nT2 > 64, nRefAngles > 64, nTE > 96) */
  DECAES_ENOMEM = -4
} decaes_status;

/* Reg (src/types.jl:66-68, dispatch src/T2mapSEcorr.jl:440-449) */
typedef enum {
  DECAES_REG_NONE = 0,
  DECAES_REG_LCURVE = 1,
  DECAES_REG_GCV = 2,
  DECAES_REG_CHI2 = 3,
  DECAES_REG_MDP = 4
} decaes_reg;

/* Flat mirror of T2mapOptions{Float64} (src/types.jl:19-100).  `nothing` => NaN. */
typedef struct {
  int32_t nx, ny, nz;      /* MatrixSize                                   */
  int32_t nTE;             /* >= 4                                          */
  int32_t nT2;             /* >= 2                                          */
  int32_t nRefAngles;      /* default 64                                    */
  int32_t nRefAnglesMin;   /* default min(5, nRefAngles)                    */
  int32_t reg;             /* decaes_reg                                    */
  int32_t legacy;          /* legacy = true algorithms (src/types.jl:20-21)  */
  int32_t alpha_provided;  /* out->alpha holds a B1 map on entry (src/T2mapSEcorr.jl:220-225) */
  int32_t ngpus;           /* host API only: 0 = all visible devices        */
  int32_t reserved;
  double TE;               /* seconds                                       */
  double T2min, T2max;     /* T2Range                                       */
  double T1;               /* default 1.0                                   */
  double Threshold;        /* default 0.0; -Inf processes every voxel       */
  double MinRefAngle;      /* default 50.0                                  */
  double RefConAngle;      /* default 180.0                                 */
  double Chi2Factor;       /* NaN when unset; required > 1 for reg=chi2     */
  double NoiseLevel;       /* NaN when unset; required > 0 for reg=mdp      */
  double SetFlipAngle;     /* NaN when unset                                */
} decaes_t2map_opts;

/* Flat mirror of T2partOptions{Float64} (src/types.jl:139-172). */
typedef struct {
  int32_t nx, ny, nz;
  int32_t nT2;
  double T2min, T2max;
  double SPWin_lo, SPWin_hi;
  double MPWin_lo, MPWin_hi;
  double Sigmoid;          /* NaN when unset */
} decaes_t2part_opts;

/* Output bundle (T2Maps / T2Distributions, src/T2mapSEcorr.jl:2-19, 64-66).
 * NULL = not requested.  Lengths are Nvox (x extra dims where noted). */
typedef struct {
  double *gdn, *ggm, *gva, *fnr, *snr, *alpha; /* required                                   */
  double *dist;                                /* required, Nvox*nT2                          */
  double *resnorm;                             /* SaveResidualNorm                            */
  double *decaycurve;                          /* SaveDecayCurve, Nvox*nTE                    */
  double *mu, *chi2factor;                     /* SaveRegParam                                */
  double *decaybasis;                          /* SaveNNLSBasis, Nvox*nTE*nT2                 */
  double *sfr, *sgm, *mfr, *mgm;               /* fused T2part outputs (need part opts)       */
} decaes_t2map_out;

/* Filled by decaes_get_stats after a host or device call (last call on this thread). */
typedef struct {
  int64_t voxels_total;     /* Nvox handed in                                   */
  int64_t voxels_processed; /* voxels above Threshold                           */
  int32_t ngpus_used;
  int32_t kernel_launches;  /* launches of our kernels during the call          */
  double setup_ms;          /* basis-set kernel (device time, max over devices) */
  double pipeline_ms;       /* voxel pipeline kernel (device time, max)         */
  double h2d_ms, d2h_ms;    /* host API only                                    */
  double total_ms;          /* host wall time of the call                       */
  /* ---- ABI v2: voxels the north_star wants "counted and reported" ---- */
  int64_t early_returns;    /* voxels that took an early-return branch of lsqnonneg_chi2!/lsqnonneg_mdp!
                               (src/lsqnonneg.jl:510-515, 708-718): the reference's save_results! reads a stale
                               cache slot there (reference-undefined); this library returns the chooser's own
                               result.  Legacy chi2 searches that ended at mu = 0 / did not end are counted too. */
  int64_t lcurve_overflow;  /* L-curve searches that outgrew the per-voxel point / state cache (the reference's
                               GrowableCache, src/utils.jl:140-254, grows without bound); 0 on every benchmark
                               configuration — a non-zero value means the result of that voxel is approximate. */
  int64_t nnls_itercap;     /* NNLS solves stopped by the 3n iteration cap (mode = 1, src/NNLS.jl:693-698)      */
  int64_t pinned_staging;   /* host API: 1 when pageable caller buffers were staged through the library's pinned
                               ring (0: the caller's buffers were already page-locked and used directly)       */
} decaes_run_stats;

/* ---- host-pointer API (what the Julia shim ccalls) ---- */
int decaes_t2map(const double *image, const decaes_t2map_opts *opts,
                 const decaes_t2part_opts *part /* NULL = no fused T2part */,
                 const decaes_t2map_out *out);

/* Same call for a Float32 volume: the image is converted to Float64 on the device (exact), everything downstream -
 * arithmetic and outputs - is Float64.  It stands for load_image's copyto!(Array{Float64,4}(undef, sz), data)
 * (src/main.jl:612-617) followed by T2mapSEcorr, with half the host-to-device traffic. */
int decaes_t2map_f32(const float *image, const decaes_t2map_opts *opts,
                     const decaes_t2part_opts *part, const decaes_t2map_out *out);

int decaes_t2part(const double *dist, const decaes_t2part_opts *part,
                  double *sfr, double *sgm, double *mfr, double *mgm);

/* Host buffers.  The host entry points accept ANY host memory.  Page-locked buffers (decaes_host_alloc,
 * cudaHostAlloc, cudaHostRegister) are copied directly; pageable ones - ordinary Julia Arrays - are staged through a
 * pinned ring the library owns (4 x 32 MB per device, DECAES_STAGE_MB), packed / unpacked by the calling host thread
 * while the neighbouring sub-slab computes.  decaes_run_stats.pinned_staging tells which path ran.
 * With more than one device and a finite Threshold the slabs are cut so that every device gets the same number of
 * voxels above Threshold (one pass over the first echo), not the same number of voxels. */
void *decaes_host_alloc(size_t bytes); /* page-locked, portable across devices; NULL on failure */
void decaes_host_free(void *p);

/* echotimes[nTE], t2times[nT2], refangleset[nRefAngles or 1],
 * decaybasisset[nTE*nT2*nRefAngles] (or nTE*nT2 with SetFlipAngle); any may be NULL. */
int decaes_setup_tables(const decaes_t2map_opts *opts, double *echotimes, double *t2times,
                        double *refangleset, double *decaybasisset);

/* ---- device-pointer API (inputs already resident in HBM) ----
 * Processes voxels [0, nvox) of arrays whose echo/bin stride is `stride` elements
 * (stride >= nvox; stride == Nvox of the full volume when working on a slab). */
int decaes_t2map_device(const double *d_image, int64_t nvox, int64_t stride,
                        const decaes_t2map_opts *opts, const decaes_t2part_opts *part,
                        const decaes_t2map_out *d_out, void *stream);

int decaes_t2part_device(const double *d_dist, int64_t nvox, int64_t stride,
                         const decaes_t2part_opts *part, double *d_sfr, double *d_sgm,
                         double *d_mfr, double *d_mgm, void *stream);

/* Synthetic MSE volume in the style of mock_image (src/utils.jl:623-658), generated on
 * the device: bi-exponential EPG signal + Rician noise, counter-based RNG keyed by
 * (seed, first_voxel + v).  d_image is [nTE][stride]. */
int decaes_mock_image_device(double *d_image, int64_t nvox, int64_t stride, int64_t first_voxel,
                             int32_t nTE, double TE, double T1, double SNR, uint64_t seed,
                             void *stream);

/* Voxel slab [*v0, *v1) owned by shard `index` of `nshards` (contiguous, boundaries aligned to the
 * 4-voxel work group; the last shard takes the remainder).  Pure host arithmetic: this is the
 * whole multi-GPU "protocol" of the path — shards are independent and there is no collective. */
int decaes_slab_bounds(int64_t nvox, int32_t nshards, int32_t index, int64_t *v0, int64_t *v1);

/* The cuts decaes_t2map uses when ngpus > 1 and Threshold is finite: cuts[0..nshards], cuts[d] .. cuts[d+1] is shard
 * d; every shard holds (to within one 1024-voxel block) the same number of voxels with first_echo[v] > threshold
 * (src/T2mapSEcorr.jl:177 is the filter).  Pure host arithmetic. */
int decaes_slab_bounds_masked(const double *first_echo, int64_t nvox, double threshold, int32_t nshards, int64_t *cuts);

/* ---- misc ---- */
const char *decaes_last_error(void);
int decaes_device_count(void);
int decaes_abi_version(void);
void decaes_get_stats(decaes_run_stats *stats);
/* Frees the device workspaces the library keeps between calls (basis tables, per-warp scratch and, for the host
 * API, the device copy of each GPU's voxel slab).  They are grow-only caches; the next call re-allocates. */
void decaes_release(void);
/* Measured DFMA peak of the current device in FLOP/s (independent FMA chains on all SMs). */
int decaes_measure_fp64_peak(double *flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* DECAES_CUDA_H */
