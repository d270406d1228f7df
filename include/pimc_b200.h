/* pimc_b200.h -- C ABI of libpimc_b200.so, the B200 (sm_100a) implementation of the
 * DelMaestroGroup/pimc measurement hot path: S(q), F(q,tau) and the per-slice pair-potential sums.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, no exceptions, no
 * exit()/assert across it.  Every entry point returns 0 on success or a negative PIMCB_E* code;
 * pimcb_last_error() returns the message of the calling thread's last failure.
 *
 * Reference interfaces each entry point replaces are cited as file:line relative to the upstream
 * tree.  The reference-side binding (EstimatorBase / LocalAction subclasses) is shown in
 * INTEGRATION.md and implemented in pimc_b200/host/.
 *
 * Call pattern (one measurement, reference `EstimatorBase::accumulate()`, src/estimator.cpp:245-252):
 *     pimcb_stage_beads(ctx, path.get_beads_data_pointer(), M, N, N_ext);   // snapshot, async H2D
 *     pimcb_ssf(ctx, sf);  pimcb_isf(ctx, isf);                             // estimator += sf, isf
 * A ctx is not thread-safe; distinct ctxs (one per device / process) are independent.
 */
#ifndef PIMC_B200_H
#define PIMC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pimcb_ctx pimcb_ctx;

enum {
    PIMCB_OK = 0,
    PIMCB_EINVAL = -1,   /* bad argument / call order                                  */
    PIMCB_ECUDA = -2,    /* a CUDA runtime call failed (message has the CUDA error)    */
    PIMCB_ENOMEM = -3,   /* host or device allocation failed                           */
    PIMCB_ESTATE = -4    /* required state missing (no box / q-vectors / beads / table) */
};

#define PIMCB_NPCFSEP 50 /* include/common.h:85, length of ActionBase::sepHist */

/* ---- lifetime --------------------------------------------------------------------------------
 * Replaces the ctor/dtor device bookkeeping of the shipped GPU estimator classes
 * (src/estimator.cpp:3751-3817, 3976-4049): streams, cudaMallocAsync'd d_beads/d_qvecs/d_out. */
int pimcb_create(pimcb_ctx** ctx, int device, int ndim);
int pimcb_destroy(pimcb_ctx* ctx);
const char* pimcb_last_error(void);
/* Library/ABI version: major*10000 + minor*100 + patch. */
int pimcb_version(void);

/* ---- geometry and wave-vectors ----------------------------------------------------------------
 * Container::side / periodic (include/container.h:24-59, src/container.cpp:84-144). */
int pimcb_set_box(pimcb_ctx* ctx, const double* side /*[ndim]*/, const unsigned* periodic /*[ndim] or NULL = all 1*/);

/* q-vectors in the AoS order produced by EstimatorBase::getQVectors (src/estimator.cpp:439-570),
 * i.e. `qValues` / `qValues_dVec` (src/estimator.cpp:3668, 3886-3892).  Each q is classified as
 * commensurate with the periodic box (q_i*side_i/2pi integer to 1e-9 in every periodic dimension,
 * zero component in non-periodic ones) or not; commensurate q take the factorised path for S(q),
 * the others the direct minimum-image pair sum. */
int pimcb_set_qvecs(pimcb_ctx* ctx, const double* q_aos /*[nq][ndim]*/, int nq);
/* Number of q-vectors classified commensurate (for diagnostics / tests). */
int pimcb_num_commensurate(const pimcb_ctx* ctx);
/* rho_q build kernel: 0 = generic (one sincos per (q, bead)); 1 = lattice path when every q is commensurate
 * (default): phase-power tables + sign-symmetry groups, particle sums as a small GEMM on the FP64 tensor cores
 * (DMMA), falling back to the CUDA-core lattice kernel and then to the generic one when the q-set does not fit;
 * 2 = force the CUDA-core lattice kernel.  Results agree to rounding; the switch exists for A/B measurement. */
int pimcb_set_rho_mode(pimcb_ctx* ctx, int mode);
/* tau-correlation kernel: FP64 tensor cores (DMMA) when M <= 510 -- 1 = one CTA per (four configurations, q), 2 =
 * persistent CTAs that stage the next pair while the current one is multiplied (bit-identical to 1) -- or 0 = CUDA-core
 * register-tiled kernel (any M).  Results agree to rounding; A/B switch. */
int pimcb_set_corr_mode(pimcb_ctx* ctx, int mode);

/* ---- bead staging ------------------------------------------------------------------------------
 * Replaces the per-call H2D of the whole padded AoS array + full sync
 * (src/estimator.cpp:3833-3842, 4074-4085).  `beads_aos` is Path::get_beads_data_pointer()
 * (include/path.h:208-210): row-major double[M][N_ext][ndim], active beads of every slice in columns
 * [0,N) (diagonal configuration, src/estimator.cpp:228).  The call snapshots the beads (the caller may
 * mutate them as soon as it returns) into the next pinned SoA buffer and enqueues the H2D copy on
 * the ctx stream; staging is shared by every estimator using the ctx. */
int pimcb_stage_beads(pimcb_ctx* ctx, const double* beads_aos, int M, int N, int N_ext);
/* B independent configurations, contiguous double[B][M][N_ext][ndim] (walker batch). */
int pimcb_stage_batch(pimcb_ctx* ctx, const double* beads_aos, int B, int M, int N, int N_ext);
/* Asynchronous form for page-locked sources (pimcb_host_alloc / pimcb_host_register): returns as soon as the DMA is
 * enqueued, so a pipelined caller keeps the host link busy back to back; `beads_aos` must stay untouched until
 * pimcb_stage_wait returns (or until results computed from it have been read).  Pageable sources are snapshotted
 * synchronously exactly as by pimcb_stage_batch. */
int pimcb_stage_batch_async(pimcb_ctx* ctx, const double* beads_aos, int B, int M, int N, int N_ext);
int pimcb_stage_wait(pimcb_ctx* ctx);
/* Same, into device slot `slot` (0 <= slot < pimcb_num_slots) without making it current; used to keep
 * several batches resident.  pimcb_select_slot makes a staged slot the current input. */
int pimcb_num_slots(const pimcb_ctx* ctx);
int pimcb_stage_batch_slot(pimcb_ctx* ctx, int slot, const double* beads_aos, int B, int M, int N, int N_ext);
int pimcb_select_slot(pimcb_ctx* ctx, int slot);
/* Pinned host memory for callers that want zero-bounce staging (beads allocated here, or an existing
 * allocation registered in place, are DMA'd directly as AoS and transposed on the device). */
int pimcb_host_alloc(void** ptr, size_t bytes);
int pimcb_host_free(void* ptr);
int pimcb_host_register(void* ptr, size_t bytes);
int pimcb_host_unregister(void* ptr);

/* ---- estimators (per staged configuration) -------------------------------------------------------
 * pimcb_ssf: replaces StaticStructureFactorEstimator::accumulate (src/estimator.cpp:3705-3737) and
 *   StaticStructureFactorGpuEstimator::accumulate (:3822-3861).  out[b][q] = sf(q)/N, the value the
 *   reference adds to `estimator` (CPU convention; norm 1/M is applied by EstimatorBase::output).
 * pimcb_isf: replaces IntermediateScatteringFunctionEstimator::accumulate (:3923-3961) and the GPU
 *   class (:4063-4100).  out[b][q*M + tau] = isf/N for tau = 0..M-1 (CPU column layout, norm 1/M).
 * Both synchronise the ctx stream before returning. */
int pimcb_ssf(pimcb_ctx* ctx, double* out /*[B][nq]*/);
int pimcb_isf(pimcb_ctx* ctx, double* out /*[B][nq*M]*/);
/* Both from one pass (rho_q is built once): either pointer may be NULL. */
int pimcb_ssf_isf(pimcb_ctx* ctx, double* ssf_out, double* isf_out);
/* pimcb_stage_beads + pimcb_ssf_isf of ONE configuration with a single synchronisation (what an estimator's
 * accumulate() needs): `beads_aos` may be mutated as soon as the call returns. */
int pimcb_ssf_isf_beads(pimcb_ctx* ctx, const double* beads_aos, int M, int N, int N_ext, double* ssf /*[nq] or NULL*/,
                        double* isf /*[nq*M] or NULL*/);

/* ---- device-resident accumulation (bins) -----------------------------------------------------------
 * `estimator += sf/N` / `estimator += isf/N` kept on the device across measurements so that only one
 * D2H (and, multi-GPU, one reduce) happens per output bin (EstimatorBase::output, src/estimator.cpp:348-362).
 * pimcb_measure enqueues rho_q build + tau-correlation (+ direct S(q) for non-commensurate q) + bin
 * accumulation of all staged configurations on the ctx stream and returns without synchronising. */
int pimcb_measure(pimcb_ctx* ctx);
int pimcb_reset_bins(pimcb_ctx* ctx);
/* Lays out the zeroed bin for M time slices before any measurement (q-vectors must be set): a rank whose share of a
 * walker batch is empty can then take part in pimcb_reduce_bins / pimcb_gather_bins_q with a zero contribution. */
int pimcb_init_bins(pimcb_ctx* ctx, int M);
/* Copies bins to the host (synchronises): ssf[nq], isf[nq*M] sums over accumulated configurations;
 * *num_accumulated = configurations in the bin. */
int pimcb_read_bins(pimcb_ctx* ctx, double* ssf /*[nq] or NULL*/, double* isf /*[nq*M] or NULL*/, long* num_accumulated);
/* Device address of the contiguous bin buffer double[nq + nq*M] (ssf then isf), for a caller-side
 * NCCL reduce over NVLink (one collective per bin).  *count = nq + nq*M. */
int pimcb_bins_device_ptr(pimcb_ctx* ctx, void** dptr, size_t* count);
int pimcb_sync(pimcb_ctx* ctx);
/* The ctx's CUDA stream (cudaStream_t as void*) so that callers can order their own work after it. */
int pimcb_stream(pimcb_ctx* ctx, void** stream);

/* ---- pair potential -------------------------------------------------------------------------------
 * Flat view of a TabulatedPotential (include/potential.h:148-157): lookupV / lookupdVdr, tableLength,
 * dr, extV, extdVdr.  Tables are copied to the device once. */
int pimcb_set_pair_table(pimcb_ctx* ctx, const double* V, const double* dVdr /*or NULL*/, int len, double dr,
                         const double* extV /*[2]*/, const double* extdVdr /*[2] or NULL*/);
/* All slices of all staged configurations in one pass.  Replaces M calls each of
 * LocalAction::V(slice) (src/action.cpp:902-947; interaction part, worm factor 1) incl. its sepHist
 * side effect (:216-224, bin = int(r/dSep)), and LocalAction::gradVSquared(slice) (:1188-1223;
 * interaction part).  vint[b][M]; f2[b][M] or NULL; sephist[b][M][50] or NULL.
 * `f2_parity`: -1 = every slice, 0/1 = only slices with slice%2 == f2_parity (others get 0). */
int pimcb_pair_sums(pimcb_ctx* ctx, double* vint, double* f2, int* sephist, double dSep, int f2_parity);
/* Non-trivial external potentials: LocalAction::gradVSquared adds externalPtr->gradV(r_i) to the pair force of every
 * bead before squaring (src/action.cpp:1216).  `gext_aos` holds those gradients for the CURRENTLY staged beads, in the
 * beads' own AoS shape ([B][M][N_ext][ndim]; evaluated by the caller through the reference's PotentialBase, O(N M));
 * it is used by pimcb_pair_sums until new beads are staged.  NULL clears it ("free" external potential, the default). */
int pimcb_set_external_gradient(pimcb_ctx* ctx, const double* gext_aos);
/* The virial terms see the external potential twice (src/action.cpp:1446-1784): gV_i = externalPtr->gradV(r_i) + sum_j
 * gradV(r_ij) in all four terms (the gradient above), and inside the T-matrix of the second-order terms dV = dVi + |gVe_i|,
 * d2V = g2Vi + externalPtr->grad2V(r_i) (:1522-1523, 1546-1547, 1721-1722).  `g2ext_aos` holds those Laplacians for the
 * CURRENTLY staged beads, one double per bead ([B][M][N_ext]); used by pimcb_virial_sums until new beads are staged.
 * NULL clears it.  With either array set, pimcb_virial_sums runs the both-ends kernel with the external terms. */
int pimcb_set_external_laplacian(pimcb_ctx* ctx, const double* g2ext_aos);

/* ---- scattering variants ---------------------------------------------------
 * pimcb_elastic: replaces ElasticScatteringEstimatorGpu::accumulate (src/estimator.cpp:4197-4235; kernel
 *   gpu_isf<true>, src/estimator_gpu.cu:66-164, launcher :408-413).  out[b][q] = the value upstream adds to
 *   `estimator` per measurement: 2/(N M) sum_{tau=0}^{M/2} sum_t sum_{i,j} cos(q.(r_j(t+tau) - r_i(t))) (its norm is
 *   0.5).  Reuses the S(q)/F(q,tau) pass of the same staged batch when pimcb_ssf_isf ran just before.
 * pimcb_ssf_cyl: replaces CylinderStaticStructureFactorEstimator::accumulate (src/estimator.cpp:5415-5456).
 *   out[b][q] = sum_t sum_{i in} [1 + 2 sum_{j>i, j in} cos(q.minimage(r_i - r_j))] with "in" = x^2 + y^2 < maxR^2
 *   (include(), :4309-4311), NOT yet divided by the particle count; n_inside[b] (may be NULL) = beads of slice 0
 *   inside the radius (num1DParticles, :4318-4327), the divisor upstream uses.  The caller sums the q of one
 *   magnitude shell (getQVectors2, :762-837).  The shipped GPU variant (gpu_ssf_cyl, src/estimator_gpu.cu:169-285)
 *   masks with the opposite sense and overwrites instead of accumulating in its strided loop; the CPU class is the
 *   semantics followed here. */
int pimcb_elastic(pimcb_ctx* ctx, double* out /*[B][nq]*/);
int pimcb_ssf_cyl(pimcb_ctx* ctx, double maxR, double* out /*[B][nq]*/, int* n_inside /*[B] or NULL*/);

/* ---- virial slice sums ------------------------------------------------------
 * Third table of the TabulatedPotential view: lookupd2Vdr2 / extd2Vdr2 (include/potential.h:148-157), same length
 * and dr as the tables given to pimcb_set_pair_table (which must come first, with dV/dr). */
int pimcb_set_pair_table_d2(pimcb_ctx* ctx, const double* d2Vdr2, int len, const double* extd2Vdr2 /*[2] or NULL*/);
/* All slices of all staged configurations in one pass: the O(N^2) parts of LocalAction::rDOTgradUterm1/2
 * (src/action.cpp:1446-1575) and deltadotgradUterm1/2 (:1588-1784), interaction potential only (external "free"):
 *   out[b][t][0] = sum_i gV_i . r_i        out[b][t][1] = sum_i (T_i gV_i) . r_i
 *   out[b][t][2] = sum_i gV_i . delta_i    out[b][t][3] = sum_i (T_i gV_i) . delta_i
 * gV_i = sum_{j != i} gradV(minimage(r_i - r_j)), T_i the per-particle T-matrix of :1547-1554.  `delta_aos`
 * ([B][M][N_ext][ndim], the beads' own AoS shape; NULL = skip, entries 2 and 3 are 0) holds each bead's deviation
 * from the centroid of its world-line window -- it depends on the link structure, which stays on the host.
 * `t2_parity`: slices that carry the T-matrix terms (gradVFactor[slice%2] > EPS): -1 all, 0 even, 1 odd, -2 none.
 * The caller applies VFactor*tau and 2*gradVFactor*tau^3*lambda. */
int pimcb_virial_sums(pimcb_ctx* ctx, const double* delta_aos, int t2_parity, double* out /*[B][M][4]*/);

/* ---- multi-GPU exchange step ------------------------------------------------------
 * One process (and one ctx) per GPU.  The path has no collective inside a measurement; per OUTPUT BIN there is one
 * NCCL reduce over NVLink (walker-configuration sharding: every rank accumulated its own configurations) or one
 * all-gather (q-vector sharding: every rank owns a contiguous range of output columns).  NCCL is resolved with dlopen
 * when the first of these is called (PIMCB_NCCL_LIB, else libnccl.so.2), so single-GPU users need no NCCL at all.
 * The reference has no multi-GPU path; these replace nothing upstream.
 *   pimcb_comm_unique_id: 128-byte ncclUniqueId created on one rank; the caller ships it to the others (MPI, a file,
 *     a torch store...).
 *   pimcb_comm_init / pimcb_comm_destroy: collective over all `nranks` ctxs.
 *   pimcb_reduce_bins: sums every rank's device bin (S(q) and F(q,tau) accumulators, layout of pimcb_read_bins) and
 *     its configuration count onto `root`, in place, on the ctx stream; afterwards pimcb_read_bins on the root
 *     returns the global bin and count (*num_total gets the count on the root and 0 elsewhere; with num_total == NULL
 *     the call does not synchronise with the device at all -- the count reaches the host with the next
 *     pimcb_read_bins).  The other ranks' bins are unchanged: call pimcb_reset_bins on every rank to start the next bin.
 *   pimcb_reduce_bins_begin / pimcb_reduce_bins_end: the same reduce, pipelined.  _begin snapshots the bin (and its
 *     count) and starts the reduce of the snapshot on the library's own communication stream, without waiting for
 *     anything; the caller resets the bin (pimcb_reset_bins) and measures the next one while the snapshot travels.  _end
 *     waits for that reduce and, on the root, copies the global bin out (ssf[nq], isf[nq][M], either may be NULL;
 *     *num_total = configurations in it, 0 on the other ranks).  One exchange in flight per ctx; every rank calls both.
 *   pimcb_gather_bins_q: nq_per_rank[nranks] = wave-vectors held by each rank (rank order = q order); every rank
 *     receives the concatenated bins ssf[sum nq] and isf[sum nq][M] (host pointers, either may be NULL). */
int pimcb_comm_unique_id(void* id_out /*[128]*/);
int pimcb_comm_init(pimcb_ctx* ctx, int nranks, int rank, const void* unique_id /*[128]*/);
int pimcb_comm_destroy(pimcb_ctx* ctx);
int pimcb_reduce_bins(pimcb_ctx* ctx, int root, long* num_total);
int pimcb_reduce_bins_begin(pimcb_ctx* ctx, int root);
int pimcb_reduce_bins_end(pimcb_ctx* ctx, double* ssf /*[nq] or NULL*/, double* isf /*[nq*M] or NULL*/, long* num_total /*or NULL*/);
int pimcb_gather_bins_q(pimcb_ctx* ctx, const int* nq_per_rank, double* ssf, double* isf);

/* ---- measurement helpers (bench) --------------------------------------------------------------------
 * Sustained FP64 DFMA throughput of the device in TFLOP/s (register-resident FMA chains on every SM,
 * timed with CUDA events).  MEASURED_PEAKS.json carries no FP64 figure. */
int pimcb_measure_fp64_peak(pimcb_ctx* ctx, double* tflops, double seconds_target);
/* Bare host-to-device rate (GB/s) of `reps` back-to-back cudaMemcpyAsync of a page-locked buffer on this context's copy
 * stream: the in-run ceiling of the host-fed (e2e) path.  Measurement only. */
/* info[5] = {(V, dV/dr) packed table in use, its verbatim (RAW) sectors, (dV/dr, d2V/dr2) packed table in use, its RAW
 * sectors, sectors per table}: the one-sector-per-pair encoding of the lookup tables (pimc_b200/csrc/table_codec.h), built
 * and verified bit for bit against the verbatim tables by pimcb_set_pair_table / pimcb_set_pair_table_d2. */
int pimcb_table_codec_info(const pimcb_ctx* ctx, long* info);
int pimcb_measure_h2d_peak(pimcb_ctx* ctx, const void* pinned_src, size_t bytes, int reps, double* gbs);
/* Per-kernel device time.  With profiling on, every kernel launch is bracketed by CUDA events on the stream it is
 * launched on; pimcb_kernel_times synchronises, folds the pending event pairs into running totals and returns, per
 * kernel id, the summed duration in ms and the number of launches since the last reset:
 * [0]=rho_q build, [1]=tau-correlation, [2]=direct S(q), [3]=bin accumulate, [4]=pair sums, [5]=AoS->SoA transpose,
 * [6]=scattering variants (elastic, cylinder S(q)), [7]=virial slice sums.
 * `on`: 0 = off, 1 = every kernel, otherwise a mask with bit (id + 1) set for each kernel id to time (an event
 * between two kernels keeps the second from being staged behind the first, so timing all of them costs ~10 % of a
 * step; the bench times only the dominant kernel inside its timed region). */
int pimcb_set_profiling(pimcb_ctx* ctx, int on);
/* Bracket only every `stride`-th launch of each profiled kernel (default 1 = every launch). */
int pimcb_set_profiling_stride(pimcb_ctx* ctx, int stride);
int pimcb_kernel_times(pimcb_ctx* ctx, double* ms_total /*[8]*/, long* count /*[8]*/, int reset);
/* Description of the rho_q build the last measurement used, for flop accounting: info[12] = {path (0 generic
 * sincos kernel, 1 DMMA lattice kernel, 2 CUDA-core lattice kernel), sign-symmetry groups, L rows, R cols, M tiles,
 * N tiles, nmax_x, nmax_y, nmax_z, commensurate q, non-commensurate q, nq}. */
int pimcb_rho_plan_info(const pimcb_ctx* ctx, int* info /*[12]*/);
/* Number of kernel launches issued by this ctx since creation. */
long pimcb_launch_count(const pimcb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PIMC_B200_H */
