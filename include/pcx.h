/* pcx.h -- C ABI of the pycollo_b200 NLP-callback engine (libpcx.so).
 *
 * Drop-in boundary for the per-iterate callbacks of pycollo's direct
 * collocation NLP.  Nothing like this exists in the reference (it is 100 %
 * Python over CasADi); each entry point names the reference interface it
 * replaces.  Paths are relative to the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types;
 *   - every function returns 0 on success, a negative PCX_E* code otherwise;
 *     pcx_last_error(engine) (or pcx_last_error(NULL) after a failed
 *     pcx_create) gives the message; no exception crosses this boundary;
 *   - x is the *scaled* iterate x_tilde (x = V*x_tilde + r), layout
 *     pycollo/backend.py:1433-1457; c layout backend.py:1551-1563;
 *   - Jacobian values follow CasADi CCS order (sorted by column, then row):
 *     pycollo/backend.py:1747-1761; Hessian values are the upper triangle in
 *     CCS order of sigma*J + lam.c (ca.nlpsol convention, backend.py:1693);
 *   - `space` says where the caller's buffers live: PCX_HOST (pageable or
 *     pinned host memory; the call copies in/out and synchronises) or
 *     PCX_DEVICE (device pointers; the call only enqueues on `stream`);
 *   - one engine = one CUDA device + one problem + one mesh; an engine is not
 *     thread-safe, distinct engines are independent;
 *   - `batch` independent iterates of the same problem are evaluated per call
 *     (multi-start / parameter sweep): every vector argument is
 *     batch-major, instance i at offset i * length.
 *   - there is NO CPU fallback: without a CUDA device every evaluation fails
 *     with PCX_ECUDA.
 */
#ifndef PCX_H
#define PCX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCX_OK        0
#define PCX_EINVAL   -1   /* bad argument / missing table                 */
#define PCX_ECUDA    -2   /* CUDA runtime or driver error                 */
#define PCX_ENVRTC   -3   /* NVRTC compilation of the kernels failed      */
#define PCX_ENOMEM   -4

#define PCX_HOST   0
#define PCX_DEVICE 1

/* evaluation selectors (bit-or) */
#define PCX_EVAL_C     1   /* constraint vector                            */
#define PCX_EVAL_DY    2   /* state derivatives at all nodes               */
#define PCX_EVAL_JAC   4   /* Jacobian non-zeros                           */
#define PCX_EVAL_HESS  8   /* Lagrangian Hessian non-zeros                 */
#define PCX_EVAL_F     16  /* objective                                    */
#define PCX_EVAL_GRAD  32  /* objective gradient (dense)                   */
/* pcx_eval_many only: the argument sets are independent -- no evaluation reads
 * or overwrites a buffer that one of the 3 preceding evaluations writes (a sweep
 * over iterates, >= 4 distinct sets).  Consecutive kernels then do not wait for
 * each other's completion and their load / compute / store phases interleave.  */
#define PCX_EVAL_INDEPENDENT 256
/* PCX_HOST evaluations of the Jacobian only: the caller's `jac` array still holds
 * what the previous host-space evaluation of this engine left in it.  Slots that do
 * not depend on the iterate (the +-1 entries of the difference operator; with fixed
 * phase times also h/2*I[l,m] entries of equations like dq/dt = qd -- whole variable
 * blocks, listed in the optional table "g_const_ranges") are then fetched only by the
 * first evaluation into that array after pcx_set_scaling; later ones copy the
 * iterate-dependent runs only (cart-pole: 20 % fewer PCIe bytes).  Passing another
 * array, or calling pcx_set_scaling, makes the next evaluation fetch everything.   */
#define PCX_EVAL_CONST_RESIDENT 512

typedef struct pcx_engine pcx_engine;

/* A named host array uploaded once at creation (sparsity maps, tiling,
 * quadrature tables ...).  Names and element types: pcx_table_name(i),
 * pcx_table_elem_size(i). */
typedef struct {
    const char* name;
    const void* data;
    int64_t     bytes;
} pcx_table;

typedef struct {
    int32_t device;          /* CUDA ordinal                                   */
    int32_t threads;         /* CTA size the tiling was built for              */
    int32_t batch;           /* instances per call                             */
    int32_t num_tiles;
    int32_t nvmax;           /* max (states+controls) over phases              */
    int32_t n_border;        /* border-map entries                             */
    int32_t bv_size;         /* border value vector length                     */
    int32_t nred_max;        /* max reductions per phase                       */
    int32_t btab_len;        /* quadrature coefficient table length            */
    int32_t reserved;        /* >0: min resident CTAs/SM (__launch_bounds__)    */
    int64_t num_x, num_c, num_dy, nnz_g, nnz_h;
    int64_t smem_bytes;      /* dynamic shared memory per CTA                  */
    const char* problem_header;  /* generated device functions (CUDA C++)     */
    int32_t num_tables;
    const pcx_table* tables;
} pcx_spec;

/* ---- lifetime ------------------------------------------------------------
 * Replaces Casadi.generate_nlp_function_callables + create_nlp_solver
 * (pycollo/backend.py:1403-1411, 1681-1693): compiles the callbacks for one
 * problem (NVRTC, sm_100a) and uploads the mesh-dependent tables.           */
int  pcx_create(const pcx_spec* spec, pcx_engine** out);
void pcx_destroy(pcx_engine* e);
const char* pcx_last_error(const pcx_engine* e);
const char* pcx_version(void);

/* Creation WITHOUT Python in the process (a compiled host: IPOPT's C++ TNLP).  The
 * problem definition is symbolic Python in the reference (OptimalControlProblem,
 * pycollo/optimal_control_problem.py) and stays there: `Engine.save_spec(path)`
 * (pycollo_b200/engine.py) writes, once per problem and mesh, everything pcx_create
 * and pcx_set_scaling are given -- the generated expression bodies, the integer
 * tables, the patterns, the scaling tables -- into one file ("PCXSPEC1": 8-byte
 * magic; int32[10] threads, batch, num_tiles, nvmax, n_border, bv_size, nred_max,
 * btab_len, reserved, num_tables; int64[6] num_x, num_c, num_dy, nnz_g, nnz_h,
 * smem_bytes; int64 header length + text; per table int64 name length + name +
 * int64 bytes + data; int64[4] lengths of pscal / gscal / border_coef / pt_scal +
 * doubles; every block padded to 8 bytes; little endian).  This call reads it,
 * creates the engine on `device` and sets the stored scaling (if any).
 * PCX_EINVAL: unreadable file, bad magic, truncated or inconsistent contents.   */
int  pcx_create_from_file(const char* path, int device, pcx_engine** out);

/* names / element sizes of the tables pcx_create expects */
int         pcx_table_count(void);
const char* pcx_table_name(int i);
int         pcx_table_elem_size(int i);

/* Replaces the substitution of numeric w, W, V, r into the CasADi graph
 * (backend.py:1459-1463, 1684-1689; scaling.py:164-170, 346-430): rewrites the
 * small scaling-dependent tables; no recompilation.                         */
int pcx_set_scaling(pcx_engine* e,
                    const double* pscal, int64_t n_pscal,
                    const double* gscal, int64_t n_gscal,
                    const double* border_coef, int64_t n_border_coef,
                    const double* pt_scal, int64_t n_pt_scal);

/* ---- evaluation ------------------------------------------------------------
 * Generic entry: any combination of PCX_EVAL_* in one fused launch.  Unused
 * outputs may be NULL.  sigma may be NULL (= 1).                            */
int pcx_eval(pcx_engine* e, int what,
             const double* x, const double* lam, const double* sigma,
             double* f, double* grad, double* c, double* dy,
             double* jac, double* hess, int space, void* stream);

/* nlp_f      : Casadi.evaluate_J             backend.py:1713-1715            */
int pcx_eval_f(pcx_engine* e, const double* x, double* f, int space, void* stream);
/* nlp_grad_f : Casadi.evaluate_g             backend.py:1717-1720            */
int pcx_eval_grad(pcx_engine* e, const double* x, double* grad, int space, void* stream);
/* nlp_g      : Casadi.evaluate_c             backend.py:1722-1725            */
int pcx_eval_c(pcx_engine* e, const double* x, double* c, int space, void* stream);
/* dy         : Casadi.dy_iter_callable       backend.py:1665-1668,
 *              solution/casadi_solution.py:71                                */
int pcx_eval_dy(pcx_engine* e, const double* x, double* dy, int space, void* stream);
/* nlp_jac_g  : Casadi.evaluate_G_nonzeros    backend.py:1738-1745            */
int pcx_eval_jac(pcx_engine* e, const double* x, double* jac, int space, void* stream);
/* nlp_hess_l : exact Hessian inside ca.nlpsol, backend.py:1693 (the
 *              reference's evaluate_H* raise NotImplementedError, :1773-1805) */
int pcx_eval_hess(pcx_engine* e, const double* x, const double* lam,
                  const double* sigma, double* hess, int space, void* stream);
/* BASELINE metric: one "eval" = Jacobian + Hessian values, one launch.       */
int pcx_eval_jac_hess(pcx_engine* e, const double* x, const double* lam,
                      const double* sigma, double* jac, double* hess,
                      int space, void* stream);

/* ---- one mesh over several GPUs (BASELINE config 4) --
 * Every rank creates the same engine (same problem, same mesh) and then
 * restricts it to a contiguous range of tiles (= a contiguous range of mesh
 * sections of each phase); x and lam are replicated.  pcx_eval then (stage 1)
 * writes only the value slots those sections own -- disjoint slabs of the
 * full-size c/dy/jac/hess arrays -- and leaves this rank's share of the few
 * cross-mesh quantities (quadrature partial sums of the integral rows and of
 * the t/s blocks of H, end-node Hessian values; nothing like it exists in the
 * reference, whose CasADi graph is one monolithic function) in a small device
 * buffer.  The caller all-reduces (sum) that buffer over the ranks (NCCL) and
 * calls pcx_apply_border (stage 2), which writes the O(1) border slots
 * (integral/endpoint rows, q/t/s corner, J, gradient entries).              */
int pcx_set_shard(pcx_engine* e, int tile_begin, int tile_end);
/* Fused variant: no NCCL call and no second launch.  The border rank allocates
 * the exchange buffer (pcx_exchange_alloc; `handle` is its CUDA IPC handle, 64
 * bytes, to be sent to the other ranks by any means), every rank attaches
 * (pcx_exchange_attach; the border rank and same-process users pass NULL / use
 * _attach_ptr).  From then on pcx_eval on a sharded engine writes the rank's
 * share of the border values straight into the border rank's memory over NVLink
 * (system-scope release), and the border rank's kernel waits for all shares,
 * sums them in rank order and writes the border slots itself.  All ranks must
 * call pcx_eval the same number of times (an epoch counter pairs the calls).
 * No host-side synchronisation between evaluations is needed: the share slots
 * are double-buffered by epoch parity and a rank reuses a slot only after the
 * border rank has published that it consumed it, so a rank runs at most two
 * evaluations ahead of the border rank.                                        */
int pcx_exchange_alloc(pcx_engine* e, int world, unsigned char handle[64]);
int pcx_exchange_attach(pcx_engine* e, int rank, int world, int border_rank,
                        const unsigned char handle[64]);
int pcx_exchange_attach_ptr(pcx_engine* e, int rank, int world, int border_rank, void* base);
int pcx_exchange_buffer(pcx_engine* e, void** base);
/* Sticky device-side status word of the engine (synchronising copy): 0 = ok,
 * 1 = a fused exchange gave up waiting for a peer rank (5 s) -- the values of
 * that evaluation are not valid.                                              */
int pcx_status(pcx_engine* e, int* code);
int pcx_shard_buffer(pcx_engine* e, double** xbuf, int64_t* n);
int pcx_apply_border(pcx_engine* e, int what, const double* x, const double* lam,
                     const double* sigma, double* f, double* grad, double* c,
                     double* jac, double* hess, void* stream);

/* ---- mesh-refinement error -----------------
 * Replaces PattersonRaoMeshRefinement.generate_dy_ph_callables +
 * phase_mesh_error (pycollo/mesh_refinement.py:88-158, 206-240) for an engine
 * that was created on the p+1 ("ph") mesh with unit scaling (V=1, r=0, W=1,
 * as mesh_refinement.py:149-150 substitutes).  x_ph is the solution
 * interpolated to the ph mesh in the engine's x layout.  One fused pass
 * evaluates dy_ph at every ph node and contracts it with the ph integration
 * blocks (the defect rows of the ph mesh ARE y_k + stretch*I*dy - Y), then a
 * section-parallel kernel forms, per phase p and section k,
 *   abs_err[err_off_p + (k*n_y + i)*mmax_p + l] = |Y_hat - Y|      (:216-226)
 *   rel_err[...] = abs / (1 + (max_l |Y| + 1))   (sic, :224, :229-234)
 *   max_rel[sec_off_p + k] = max over states and nodes            (:237-240)
 * with mmax_p = max_k N_k(ph) - 1 and zeros in the unused tail, phases
 * concatenated.  Any output may be NULL.
 * Multi-GPU: on an engine restricted to a tile
 * range (pcx_set_shard) only the sections of that range are evaluated and only
 * their entries of abs_err / rel_err / max_rel are written -- the pass is
 * section-local, there is no exchange; a global maximum, if wanted, is one
 * all_reduce(max) of max_rel over the ranks.                                  */
int pcx_mesh_error(pcx_engine* e, const double* x_ph, double* abs_err,
                   double* rel_err, double* max_rel, int space, void* stream);
/* lengths (per instance) of abs_err/rel_err and of max_rel                    */
int pcx_mesh_error_sizes(const pcx_engine* e, int64_t* n_err, int64_t* n_sections);

/* ---- solution re-fit onto the p+1 mesh ---
 * Replaces SolutionABC.interpolate_solution_{lobatto,radau}
 * (pycollo/solution/solution_abc.py:60-142: per state and section a Legendre fit
 * of dy*T/2 integrated from the section's first value; per control a polynomial
 * fit) + PattersonRaoMeshRefinement.construct_x_ph / eval_polynomials
 * (mesh_refinement.py:160-204) for an engine on the ITERATION mesh: from the
 * solution x (user basis, engine x layout) and dy (pcx_eval_dy) it writes x_ph,
 * the solution on the p+1 mesh in the x layout of the ph-mesh engine
 * (section-boundary values copied, interior ph nodes from the per-section
 * interpolants, q / t / s copied), ready for pcx_mesh_error.  The fits are applied
 * as precomputed per-order matrices (exact interpolants; the reference's
 * numpy least-squares fits in the [0,1] window lose ~1e-12 at order 10).      */
int pcx_refit_to_ph(pcx_engine* e, const double* x_user, const double* dy,
                    double* x_ph, int space, void* stream);
int pcx_refit_size(const pcx_engine* e, int64_t* num_x_ph);

/* ---- guess interpolation to a new mesh -----
 * Replaces Iteration.interpolate_guess_to_mesh (pycollo/iteration.py:86-194:
 * scipy interp1d, linear, fill_value="extrapolate", per state / control row) for
 * an engine on the NEW mesh: x_prev is a solution or guess (user basis) in the x
 * layout of a previous mesh with prev_N[p] nodes in phase p at abscissae
 * prev_tau (phases concatenated), tau the new mesh's abscissae (phases
 * concatenated, N_p each).  y and u rows are interpolated with interp1d's own
 * formula ((t - t_lo)/(t_hi - t_lo))*y_hi + ((t_hi - t)/(t_hi - t_lo))*y_lo (scipy >= 1.10), lo/hi from a left bisection clipped to
 * [1, M-1]; q, t and s are copied (iteration.py:166-168).  prev_N is a host
 * array of num_phases entries whatever `space` is.                            */
int pcx_interp_guess(pcx_engine* e, const double* x_prev, const double* prev_tau,
                     const int64_t* prev_N, const double* tau, double* x_guess,
                     int space, void* stream);

/* ---- sparsity structure for a C/C++ host ----------------------------------
 * Replaces Casadi.evaluate_G_structure (pycollo/backend.py:1747-1761) and the
 * cyipopt contract jacobianstructure() / hessianstructure()
 * (pycollo/nlp.py:36-76).  The patterns are fixed per (problem, mesh); they are
 * handed to pcx_create as the optional int64 tables "g_rows", "g_cols",
 * "h_rows", "h_cols" (value order of pcx_eval) and kept on the host.
 *   PCX_ORDER_NATIVE   : the order pcx_eval writes -- Jacobian in CCS order
 *                        (by column, then row), Hessian upper triangle in CCS
 *                        order (row <= col); perm = identity;
 *   PCX_ORDER_ROW_MAJOR: Jacobian by row, then column; Hessian as the LOWER
 *                        triangle by row, then column (the transposed upper
 *                        triangle: same value order, rows/cols swapped).
 * rows/cols/perm have nnz entries (pcx_sizes); perm may be NULL.  Values in the
 * requested order are values_native[perm[i]] (pcx_gather does it on the device). */
#define PCX_ORDER_NATIVE    0
#define PCX_ORDER_ROW_MAJOR 1
int pcx_structure_jac(const pcx_engine* e, int order, int64_t* rows, int64_t* cols,
                      int64_t* perm);
int pcx_structure_hess(const pcx_engine* e, int order, int64_t* rows, int64_t* cols,
                       int64_t* perm);

/* ---- iteration scaling from the sparse Jacobian --
 * Replaces the dense np.array(G) + row norms of
 * IterationScaling._calculate_constraint_scaling (pycollo/scaling.py:392-396):
 * evaluates the Jacobian at x with the scaling currently set and writes
 * norms[i] = sqrt(sum_j G[i,j]^2) for every constraint row (num_c doubles); the
 * Jacobian values never leave the device.  Needs the structure tables.        */
int pcx_jac_row_norms(pcx_engine* e, const double* x, double* norms, int space,
                      void* stream);

/* ---- variable / constraint bounds on the mesh ----
 * Replaces Iteration.generate_variable_bounds / generate_constraint_bounds /
 * scale_bounds (pycollo/iteration.py:408-453): OCP-level bounds (one (lo, hi)
 * pair per OCP variable in the order [per phase: y, u, q, t] ++ [s]; per OCP
 * constraint in the order [per phase: defect, path, integral] ++ [endpoint];
 * state bounds at t0 / tF per phase state) are expanded to the mesh and scaled:
 *   x_lo[i] = (1/V[i]) * (lo - r[i])   (scaling.py:172-174),  c_lo[i] = W[i] * lo.
 * All inputs are HOST arrays (they are O(#OCP variables)); outputs follow
 * `space`.  y_t0 / y_tF hold 2 doubles (lo, hi) per state, phases concatenated. */
int pcx_expand_bounds(pcx_engine* e, const double* ocp_x_bnd, const double* y_t0_bnd,
                      const double* y_tF_bnd, const double* ocp_c_bnd,
                      const double* V_ocp, const double* r_ocp, const double* W_ocp,
                      double* x_lo, double* x_hi, double* c_lo, double* c_hi,
                      int space, void* stream);

/* ---- many evaluations, one foreign call -------------------------------------
 * `count` evaluations enqueued back to back on `stream` from C, cycling through
 * `n_sets` argument sets (device pointers; a sweep over iterates, or a ring of
 * buffers larger than L2).  With gate != 0 the stream is held by a one-thread
 * kernel until every launch has been enqueued, so the device-side time does not
 * depend on how fast the host enqueues.  If elapsed_ms is not NULL the call
 * brackets the `count` launches with CUDA events on `stream` (after the gate),
 * synchronises and returns the elapsed time.  `warm` launches are enqueued first,
 * in the same call and outside the timed bracket: on a mesh sharded over several
 * GPUs their in-kernel exchange brings the ranks into step, so the timed launches
 * do not contain the ranks' start-up skew.                                    */
typedef struct {
    const double *x, *lam, *sigma;
    double *f, *grad, *c, *dy, *jac, *hess;
} pcx_args;
int pcx_eval_many(pcx_engine* e, int what, const pcx_args* sets, int n_sets,
                  int count, int warm, void* stream, int gate, float* elapsed_ms);

/* Host-space sweep over iterates (multi-start / parameter sweep with a host
 * consumer): `count` evaluations of PCX_EVAL_JAC or PCX_EVAL_JAC | PCX_EVAL_HESS
 * cycling through `n_sets` >= 2 HOST argument sets (pinned memory for full PCIe
 * speed), pipelined -- the upload of evaluation i+1 and the kernel of evaluation i
 * run under the download of evaluation i-1.  Returns when every result is in its
 * host arrays.  PCX_EVAL_CONST_RESIDENT applies per host jac array.             */
int pcx_sweep_host(pcx_engine* e, int what, const pcx_args* sets, int n_sets, int count,
                   void* stream);

/* Sizes (per instance) -- Casadi.evaluate_G_num_nonzero backend.py:1763-1771 */
int pcx_sizes(const pcx_engine* e, int64_t* num_x, int64_t* num_c, int64_t* num_dy,
              int64_t* nnz_jac, int64_t* nnz_hess, int32_t* batch);

/* Permuted views for a cyipopt-style host (row-major Jacobian, lower-triangular
 * Hessian; pycollo/nlp.py:36-76, iteration.py:930-933, 965-968):
 * out[i] = in[perm[i]] on the device.                                        */
int pcx_gather(pcx_engine* e, const double* in, const int64_t* perm, int64_t n,
               double* out, int space, void* stream);

/* Pinned host buffers for the caller (async H2D/D2H of iterates and values). */
int pcx_host_alloc(void** ptr, int64_t bytes);
int pcx_host_free(void* ptr);
/* Page-lock memory the CALLER owns (the x / lambda vectors an NLP solver hands to
 * its callbacks -- IPOPT's TNLP::eval_* `const Number* x`, pycollo/nlp.py:40-62 --
 * are the same few buffers for the whole solve), so that a PCX_HOST evaluation
 * or the caller's own cudaMemcpyAsync reads them by DMA without a staging copy.
 * PCX_ECUDA if the range cannot be registered (e.g. it already is): the buffer
 * then simply stays pageable.                                                  */
int pcx_host_register(void* ptr, int64_t bytes);
int pcx_host_unregister(void* ptr);

/* Number of kernel launches issued by this engine since creation, and the
 * names of the compiled kernel variants (diagnostics for bench.py).          */
int64_t pcx_launch_count(const pcx_engine* e);
/* Jacobian + Hessian bytes the last PCX_HOST evaluation copied device -> host      */
int64_t pcx_last_d2h_bytes(const pcx_engine* e);
/* Compiled-kernel facts of the variant for `what` (compiling it if needed): CTAs
 * resident per SM at the engine's CTA size and shared memory, registers per
 * thread, local-memory (spill) bytes per thread, static shared memory.        */
int     pcx_variant_info(pcx_engine* e, int what, int* blocks_per_sm, int* registers,
                         int* local_bytes, int* static_smem_bytes);
int     pcx_synchronize(pcx_engine* e, void* stream);
/* debug: copy the first n doubles of the reduction scratch to the host         */
int     pcx_debug_read_partials(pcx_engine* e, double* dst, int64_t n);

/* Write `bytes` bytes of scratch on the device (> L2) so the next timed launch
 * starts from a cold L2.                                                     */
int pcx_flush_l2(pcx_engine* e, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PCX_H */
