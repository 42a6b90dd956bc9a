/*
 * skeelberzins_b200.h -- C ABI of the B200-native batched Skeel-Berzins solve path.
 *
 * This is the drop-in boundary for the internal implicit-Euler / stationary Newton solvers of
 * SkeelBerzins.jl (reference paths are cited per entry point as file:line relative to the
 * reference tree).  The host (Julia via ccall, Python via ctypes, or C++) owns the time loop
 * (reference src/pdepe.jl:90-108) and the Newton loop (src/utils.jl:311-336) and calls the
 * entry points below once per Newton iteration.  All work runs in hand-written sm_100a CUDA
 * kernels; there is NO CPU fallback: every compute entry point fails with SB_ERR_CUDA when no
 * device is present.
 *
 * Conventions
 *   - every function returns an int status: SB_OK (0) or a negative SB_ERR_* code; the text of the
 *     last error of the calling thread is available from sb_last_error().  Nothing throws.
 *   - host arrays in the REFERENCE layout: one instance after another, inside an instance node by
 *     node with the npde components fastest (src/utils.jl:632, src/assembler.jl:21-22):
 *         u_host[(k*Nx + i)*npde + j]      k = instance, i = mesh node, j = component
 *   - one sb_ctx per GPU; a context is not thread-safe, distinct contexts are independent.
 *   - calls are asynchronous on the context's stream unless they move data to/from the host.
 */
#ifndef SKEELBERZINS_B200_H
#define SKEELBERZINS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_INVALID (-1)      /* bad argument (message says which)                              */
#define SB_ERR_CUDA (-2)         /* CUDA runtime error / no device                                  */
#define SB_ERR_UNSUPPORTED (-3)  /* valid request outside the kernels' compiled range               */
#define SB_ERR_CONVERGENCE (-4)  /* reference: throw("convergence failed") src/utils.jl:336,377     */
#define SB_ERR_PLUGIN (-5)       /* NVRTC compilation of a user functor failed                      */

/* ---- coefficient-functor registry ------------------------------------------------------------
 * Device-compiled restatements of the reference's example callbacks pdefun(x,t,u,dudx)->(c,f,s) and
 * bdfun(xl,ul,xr,ur,t)->(pl,ql,pr,qr).  Each instance k of a batch carries its own parameter
 * vector prm[0..nparams) (the "parameter sweep").  The values in brackets reproduce the example.
 */
enum sb_functor_id {
    /* examples/Example101_LinearDiffusion.jl:21-38 ; test/test_solveTransient.jl:7-24
     * c=1, f=D*ux, s=0 ; p=0,q=1 both ends.                      prm = {D [1]}                    */
    SB_LIN_DIFF = 0,
    /* examples/Example102_NonlinearDiffusion.jl:44-65
     * c=1, f=a*u*ux, s=0 ; p=0,q=1 both ends.                    prm = {a [2]}                    */
    SB_NONLIN_DIFF = 1,
    /* examples/Example103_LinearDiffusionCylindrical.jl:35-58  (m=1, singular)
     * c=1, f=D*ux, s=0 ; left (0,0) ; right p=ur-g*exp(-D*n^2*t), q=0.
     *                                                            prm = {D [1], n, g [=J0(n)]}      */
    SB_LIN_DIFF_CYL = 2,
    /* examples/Example104_Poisson.jl:20-41 ; test/test_solveStationary.jl:6-23 (c0=0)
     * c=c0, f=ux, s=s0 ; p=u-ub, q=0 both ends.                  prm = {c0 [1], s0 [1], ub [0.1]}  */
    SB_POISSON = 3,
    /* examples/Example105_StationaryNonlinearDiffusion.jl:30-51
     * c=1, f=a*u*ux, s=s0 ; p=u-ub, q=0 both ends.               prm = {a [2], s0 [1], ub [0.1]}   */
    SB_STAT_NONLIN = 4,
    /* examples/Example106_SystemReactionDiffusion.jl:42-64   (npde = 2)
     * c=(1,1), f=(d1,d2).*ux, y=u1-u2, s=(-k*y, k*y) ;
     * left p=(u1-1,0) q=(0,1) ; right p=(0,u2) q=(1,0).          prm = {d1 [.5], d2 [.1], k [1]}   */
    SB_REACT_DIFF = 5,
    /* examples/Example107_LinearDiffusionSpherical.jl:34-55  (m=2, singular)
     * c=1, f=D*ux, s=0 ; left (0,0) ; right p=ur-(1+6*D*t), q=0. prm = {D [1]}                     */
    SB_LIN_DIFF_SPH = 6,
    /* examples/Example201_PartialDerivativeApprox.jl:34-51
     * c=1, f=D*ux, s=-(D*eta/L)*ux ; p=u, q=0 both ends.         prm = {D [.1], eta [10], L [1]}   */
    SB_CONV_DIFF = 7,
    /* examples/Example301_PDEOptimSciMLSensitivity.jl:31-48
     * c=1, f=a1*ux, s=2*a0*u ; p=u, q=0 both ends.               prm = {a0, a1}                    */
    SB_LIN_REACT = 8,
    /* quasi-linear scalar family (covers all scalar examples and the manufactured tests):
     *   c = c0 + c1*u ; f = (d0 + d1*u)*ux ; s = s0 + s1*u + s2*ux
     *   left : p = al*ul - (gl0 + gl1*t + gl2*exp(-ll*t)), q = ql
     *   right: p = ar*ur - (gr0 + gr1*t + gr2*exp(-lr*t)), q = qr
     * prm = {c0,c1,d0,d1,s0,s1,s2, al,gl0,gl1,gl2,ll,ql, ar,gr0,gr1,gr2,lr,qr}   (19 values)       */
    SB_GENERIC_SCALAR = 9,
    SB_NUM_BUILTIN_FUNCTORS = 10,
    /* ids >= SB_USER_FUNCTOR_BASE are handed out by sb_register_user_functor()                     */
    SB_USER_FUNCTOR_BASE = 1000
};

/* registry functors have npde <= 2; user functors (plugin slot) up to 4: npde = 3 on meshes up to 1024 nodes, npde = 4 up
 * to 512 (the factors of a block row are 2*npde^2 + npde doubles per thread: rows per thread 16 / 8 / 4 / 2 for npde 1..4) */
#define SB_MAX_NPDE 4
#define SB_GENERIC_SCALAR_NPARAMS 19

typedef struct sb_ctx sb_ctx;
typedef struct sb_problem sb_problem;

/* which Newton-iteration pipeline sb_newton_iteration() runs */
enum sb_design {
    SB_DESIGN_SPLIT = 0, /* assemble kernel writes F and the block-tridiagonal J to HBM; solve kernel
                            reads them (the two calls sb_residual_jacobian + sb_solve_update_norm)  */
    SB_DESIGN_FUSED = 1, /* one kernel per iteration; J and F never leave the SM                    */
    SB_DESIGN_FUSED_TMA = 2, /* fused, persistent CTAs: u and b of the next instance are prefetched by the
                            bulk-copy engine (cp.async.bulk + mbarrier) while the current one is solved;
                            only instances still active are visited (compacted list)                */
    SB_DESIGN_RESIDENT = 3 /* default.  Per-iteration entry points (sb_newton_iteration ...) as fused_tma; the loops
                            driven from C (sb_newton_solve, sb_solve_steps, sb_solve_steps_opt) run as ONE launch in
                            which every instance stays on its SM through all Newton iterations (and time steps)  */
};

/* ---- library / context ---------------------------------------------------------------------- */
const char* sb_version(void);
/* digest of the sources (csrc/, include/) and compiler flags the library was built from; the host mirror
 * compares it with the tree it runs from, so a stale library fails loudly instead of running old kernels */
const char* sb_build_digest(void);
const char* sb_last_error(void);
int sb_device_count(int* count);
/* Guard mode (environment SB_GUARD=1 before the first call): every device array of a problem is allocated with
 * 4 KiB of canary bytes on both sides, checked when it is freed.  violations = arrays whose canaries were overwritten
 * so far, arrays_checked (may be NULL) = arrays checked so far.  A substitute for a memory checker where none can run. */
int sb_debug_guard_violations(int64_t* violations, int64_t* arrays_checked);

/* number of CUDA kernels this library has launched so far in the process (bench.py's gpu_launches) */
int sb_kernel_launches(int64_t* count);
/* stream == NULL: the context creates its own non-blocking stream; otherwise a cudaStream_t of the
 * caller (e.g. the torch current stream, so that the caller's events bracket the kernels).       */
int sb_ctx_create(int device, void* stream, sb_ctx** out);
int sb_ctx_destroy(sb_ctx* ctx);
int sb_ctx_synchronize(sb_ctx* ctx);
/* page-locked host memory for upload/download buffers (the copies of sb_state_upload/download run at
 * full PCIe/NVLink-C2C speed only from pinned memory)                                             */
int sb_host_alloc(size_t bytes, void** out);
int sb_host_free(void* p);

/* functor metadata: npde and number of per-instance parameters of a registry entry              */
int sb_functor_info(int functor_id, int* npde, int* nparams, const char** name);

/* Plugin slot: compile (NVRTC, sm_100a) a user functor given as CUDA C++ source that defines
 *   struct <struct_name> { static constexpr int npde, nparams;
 *       template<class S> static __device__ void pde(double x,double t,const S* u,const S* ux,
 *                                                    const double* prm, S* c, S* f, S* s);
 *       template<class S> static __device__ void bc (double xl,const S* ul,double xr,const S* ur,
 *                                                    double t,const double* prm,
 *                                                    S* pl,double* ql,S* pr,double* qr); };
 * i.e. the reference's pdefun/bdfun (docs/src/solvers.md:9-47) generic in the scalar type so the
 * dual numbers flow through.  Returns the new functor id in *functor_id.                          */
int sb_register_user_functor(sb_ctx* ctx, const char* cuda_source, const char* struct_name,
                             int npde, int nparams, int* functor_id);

/* NVRTC-only dry run (needs no GPU): compile the kernels a problem with Nx mesh points and coordinate symmetry m would
 * need for a registered user functor; SB_ERR_PLUGIN + the compiler log on failure, size of the CUBIN on success.  */
int sb_user_functor_compile_check(int functor_id, int Nx, int m, size_t* cubin_bytes);

/* ---- problem (replaces problem_init, src/utils.jl:609-663, for a batch of B instances) -------
 * m          coordinate symmetry 0/1/2 (asserts of src/pdepe.jl:55-57 apply)
 * xmesh_host Nx mesh points shared by the batch
 * params_host[B*nparams] per-instance functor parameters, instance-major
 * Computes singular (src/utils.jl:616), xi/zeta (src/utils.jl:683-718) and the interpolation and
 * flux/volume weights used by assemble! (src/utils.jl:18-48, src/assembler.jl:36,68-69,98).      */
int sb_problem_create(sb_ctx* ctx, int m, int Nx, int64_t B, const double* xmesh_host,
                      int functor_id, const double* params_host, int nparams, sb_problem** out);
int sb_problem_destroy(sb_problem* pb);
/* queries: npde, Nx, B, singular flag, xi/zeta (host copies, length Nx-1 each; may be NULL)      */
int sb_problem_info(sb_problem* pb, int* npde, int* Nx, int64_t* B, int* singular);
int sb_problem_quadrature(sb_problem* pb, double* xi_host, double* zeta_host);
int sb_problem_set_design(sb_problem* pb, int design);
/* replace the per-instance functor parameters [B*nparams] of an existing problem (same batch, same
 * mesh): lets a host re-use the device allocations across solves of a parameter study.            */
int sb_problem_set_params(sb_problem* pb, const double* params_host);
/* thread decomposition chosen for the batch: R rows per thread, T threads per system (T == 32: one warp
 * per system, shuffles only; T > 32: one CTA per system, shared-memory exchange)                      */
int sb_problem_layout(sb_problem* pb, int* R, int* T, int* warp);

/* state u (reference layout on the host side; the device keeps a chunk-transposed SoA layout)   */
int sb_state_upload(sb_problem* pb, const double* u_host);
int sb_state_download(sb_problem* pb, double* u_host);

/* device-side copy of the state (device layout) and its restore; rollback also clears the iteration
 * counters.  Lets a caller re-run a solve from HBM-resident inputs (bench.py's `value` leg).      */
int sb_state_checkpoint(sb_problem* pb);
int sb_state_rollback(sb_problem* pb);

/* mass_matrix (src/assembler.jl:138-174): 0/1 flags per (instance, component, node class) from the
 * CURRENT state at time t0 (call right after uploading the initial condition).  stationary != 0
 * selects implicitEuler_stat! semantics (src/utils.jl:264-276): all flags 1.                     */
int sb_mass_flags(sb_problem* pb, double t0, int stationary);
/* download flags as doubles in reference layout [B][Nx][npde] (diag of the mass matrix)          */
int sb_mass_download(sb_problem* pb, double* mass_host);

/* assemble!(du,u,pb,t) (src/assembler.jl:19-125) on reference-layout HOST vectors of B instances;
 * copies in, runs the residual-only kernel, copies out.  u_host==NULL: use the resident state.   */
int sb_assemble(sb_problem* pb, const double* u_host, double t, double* du_host);
/* The same operator on DEVICE memory -- the right-hand side f(du,u,p,t) = assemble!(du,u,pb,t) an ODE/DAE
 * integrator calls when it keeps its vectors on the GPU (src/diffEq.jl:18-27).  d_u, d_du: device pointers,
 * reference layout [B][Nx][npde]; d_u == NULL: the resident state.  Asynchronous on the context's stream: no
 * allocation after the first call, no host synchronisation.                                           */
int sb_assemble_device(sb_problem* pb, const double* d_u, double t, double* d_du);

/* pdeval(m, xmesh, u_approx, x_eval, pb) (src/utils.jl:396-438; the same evaluation serves interpolate_sol_space,
 * :460-559): value and x-derivative of the solution at nq points shared by the batch, by the interpolation the
 * discretisation itself uses (src/utils.jl:18-48).  u_host: states in reference layout, or NULL for the resident
 * state.  Outputs [B][nq][npde].  A point outside the mesh -> SB_ERR_INVALID, "Evaluation point oustide of the
 * spatial discretization." (the reference's message, :413).                                                      */
int sb_pdeval(sb_problem* pb, const double* u_host, int nq, const double* xq_host, double* u_out, double* dudx_out);

/* time step n -> n+1 (src/pdepe.jl:94, src/utils.jl:309): b = M.*u_n ; u <- b ; all instances
 * active ; per-step iteration counters cleared.  inv_tau = 1/tau (0 for the stationary solve).   */
int sb_begin_step(sb_problem* pb, double t_next, double inv_tau);

/* implicitEuler! + forwarddiff_color_jacobian! + value! + rhs update
 * (src/utils.jl:245-257, 312-314): F = -assemble(u,t) + inv_tau*M.*u - inv_tau*b and the
 * block-tridiagonal dF/du, for active instances, into device scratch.                           */
int sb_residual_jacobian(sb_problem* pb);
/* solve + update + norm + stop test (src/utils.jl:317-333): h = J\F ; u -= h ; nm = ||h||_2 ;
 * iters++ ; active &= !(nm < tol).                                                               */
int sb_solve_update_norm(sb_problem* pb, double tol);
/* one whole Newton iteration with the configured design (split = the two calls above)           */
int sb_newton_iteration(sb_problem* pb, double tol);

/* number of instances still iterating after the most recent iteration that has completed.
 * blocking != 0 waits for everything enqueued so far; otherwise returns the newest value already
 * delivered (never less than the true current value -> safe for "keep iterating" decisions).     */
int sb_poll_active(sb_problem* pb, int blocking, int64_t* n_active);
/* wait for Newton iteration `it` (0-based since sb_begin_step) and return the number of instances
 * that were still active after it; later iterations may already be enqueued behind it.           */
int sb_wait_iteration(sb_problem* pb, int it, int64_t* n_active);

/* convenience: the Newton loop of src/utils.jl:311-336 driven from C with software pipelining
 * (iterations are enqueued ahead of the poll; converged instances are frozen bit-exactly).
 * Returns SB_ERR_CONVERGENCE if any instance is still active after maxit iterations.
 * iters_done (may be NULL) = iterations launched.                                                */
int sb_newton_solve(sb_problem* pb, double tol, int maxit, int* iters_done);

/* convenience: nsteps implicit-Euler steps (the time loop of src/pdepe.jl:90-108 driven from C; no state is
 * stored in between).  tgrid[nsteps+1] is the time grid; tau > 0: constant time step, else tau_i = diff(tgrid).
 * With the resident design (default) on a mesh in the batched range this is ONE kernel launch (csrc/sb_resident.cuh): the
 * instances are independent, so every thread group keeps its instance on the SM through all Newton iterations of all
 * time steps -- same arithmetic per iteration, hence the same states and per-instance iteration counts as the
 * host-owned loops, without a launch per iteration or lock-step waiting.  The fused_tma design keeps the previous
 * scheme (one launch per Newton iteration, time stepping on the device; SB_NO_DEVICE_STEPPING=1: the plain
 * step-by-step loop).  sb_newton_solve uses the same kernel for the Newton loop of a single step.  steps_done (may be NULL) =
 * completed time steps; SB_ERR_CONVERGENCE as in sb_newton_solve.                                                    */
int sb_solve_steps(sb_problem* pb, int nsteps, const double* tgrid, double tau, double tol, int maxit, int* steps_done);

/* The same time loop with the per-instance extras of the whole-solve kernel (csrc/sb_resident.cuh: every instance stays
 * on its SM from the first to the last Newton iteration of the solve; one launch):
 *   tgrid_per_instance  host [B][nsteps+1] or NULL: every instance steps through its own time grid -- the reference's
 *                       vector tstep (src/pdepe.jl:84-85,92), one vector per instance of the batch; tgrid/tau are then ignored
 *   step_iters          host [B][nsteps] or NULL: Newton iterations of every instance in every step (src/utils.jl:323-333)
 *   snapshots           host [nsteps+1][B][Nx][npde] or NULL (pinned memory recommended): every state including the
 *                       initial one, what the reference stores (src/pdepe.jl:99-107); finished rounds of instances are
 *                       copied out on a second stream while later rounds compute
 * Needs the resident design (the default) and a mesh in the batched range; otherwise SB_ERR_UNSUPPORTED.           */
int sb_solve_steps_opt(sb_problem* pb, int nsteps, const double* tgrid, const double* tgrid_per_instance, double tau,
                       double tol, int maxit, int32_t* step_iters, double* snapshots, int* steps_done);

/* The same time loop when only u(t_end) is wanted (src/pdepe.jl:112-120 with the last state): final_state is host
 * [B][Nx][npde] (pinned memory recommended) and receives the result round by round of instances, on a second stream,
 * while later rounds still compute -- the download of sb_state_download hidden behind the solve.  The device state is
 * u(t_end) as after sb_solve_steps.  Resident design and batched range only, otherwise SB_ERR_UNSUPPORTED.          */
int sb_solve_steps_final(sb_problem* pb, int nsteps, const double* tgrid, double tau, double tol, int maxit, double* final_state,
                         int* steps_done);

/* The whole data path of a transient pdepe call (src/pdepe.jl:66-110) as one pipelined operation: initial states from
 * host memory u0_host [B][Nx][npde], mass flags at t0, the time loop, u(t_end) to host memory final_state [B][Nx][npde]
 * (both pinned for real overlap).  Equivalent to sb_state_upload + sb_mass_flags(t0, 0) + sb_solve_steps_final, except
 * that the states are uploaded in chunks of instances WHILE the whole-solve kernel already iterates on the chunks that
 * have arrived (it waits per instance on an arrival word written by the copy engine), as finished rounds are downloaded
 * while later ones iterate: neither transfer is exposed.  Resident design, batched range.                         */
int sb_solve_from_host(sb_problem* pb, const double* u0_host, double t0, int nsteps, const double* tgrid, double tau, double tol,
                       int maxit, double* final_state, int* steps_done);

/* diagnostics (host arrays of length B unless stated)                                            */
int sb_get_step_iters(sb_problem* pb, int32_t* iters_host);       /* Newton iterations this step  */
int sb_get_total_iters(sb_problem* pb, int64_t* total_host);      /* since sb_state_upload        */
int sb_get_last_norm(sb_problem* pb, double* norm_host);          /* ||h||_2 of last iteration    */
int sb_get_status(sb_problem* pb, int32_t* status_host);          /* 0 ok, 1 not converged, 2 non-finite / singular,
                                                                     3 ok, solved with the pivoted fallback      */
/* ||h||_2 history of the current step (Params.hist, src/utils.jl:305-307,325-327): enable with a
 * capacity (>= maxit), then after a step hist_host[k*cap + it] holds the norm of iteration it.     */
int sb_enable_history(sb_problem* pb, int cap);
int sb_get_history(sb_problem* pb, double* hist_host, int cap);
/* sum over instances of Newton iterations since upload (the work counter of the metric)         */
int sb_total_active_iterations(sb_problem* pb, int64_t* total);

/* Profiling window: between begin and end every Newton-iteration launch is bracketed by timed CUDA
 * events on the context's stream.  end() returns the summed kernel time (ms) of the launches that
 * had work (split: assemble and solve separately; fused: ms_assemble = 0), their number, the number
 * of no-op launches (every instance already converged) and the active instance-iterations executed. */
int sb_profile_begin(sb_problem* pb);
int sb_profile_end(sb_problem* pb, double* ms_assemble, double* ms_solve, int64_t* launches,
                   int64_t* noop_launches, int64_t* instance_iterations);

/* device-side scratch access for parity tests of the assemble kernel: F and J of the last
 * sb_residual_jacobian in reference layout: F[B][Nx][npde], Jl/Jd/Ju[B][Nx][npde][npde]
 * (row component major): Jl = d F_i / d u_{i-1}, Jd = d F_i / d u_i, Ju = d F_i / d u_{i+1}.     */
int sb_debug_download_system(sb_problem* pb, double* F, double* Jl, double* Jd, double* Ju);
/* solve given block-tridiagonal systems (reference layout as above) with the product kernels;
 * x[B][Nx][npde].  Used by the solver parity tests.                                              */
int sb_debug_solve(sb_problem* pb, const double* Jl, const double* Jd, const double* Ju,
                   const double* F, double* x);

#ifdef __cplusplus
}
#endif
#endif /* SKEELBERZINS_B200_H */
