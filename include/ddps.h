/*
 * ddps.h -- C ABI of libddps_b200.so, the B200-native (sm_100a, FP64) replacement for the hot path of
 * DriftDiffusionPoissonSystems.jl: per-iteration P1-triangle FEM assembly + linear solve inside
 * solve_ddpe (Gummel) and solve_semlin_poisson (Newton).
 *
 * The reference is a single Julia file with NO FFI/plugin interface; "src:N" below means
 * /root/reference/src/DriftDiffusionPoissonSystems.jl line N and names the reference code each entry
 * point replaces.  A Julia (ccall) or Python (ctypes) host shim evaluates the user's spatial
 * closures once on the host, hands the sampled arrays to this library, and calls the solve entry
 * points; all per-iteration work stays resident on the device.  See INTEGRATION.md for the
 * reference-side binding.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; ddps_last_error(ctx) gives the message
 *     (ctx may be NULL for errors of ddps_ctx_create).  No C++ exception crosses this boundary.
 *   - the caller owns every host pointer; the library copies during the call and never retains it.
 *   - index arrays are int64 and 1-BASED exactly like the reference's Mesh (src:9-18).
 *   - there is NO CPU fallback: without a CUDA device ddps_ctx_create fails.
 *   - calls are blocking; one ctx per host thread; one ctx (= one process) per GPU for multi-GPU.
 */
#ifndef DDPS_H
#define DDPS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDPS_VERSION 100 /* 0.1.0 */

typedef struct ddps_ctx ddps_ctx;

/* ---- error codes ------------------------------------------------------------------------- */
#define DDPS_OK 0
#define DDPS_ERR_ARG (-1)       /* bad argument / inconsistent input (reference: error()/@assert) */
#define DDPS_ERR_CUDA (-2)      /* CUDA runtime failure */
#define DDPS_ERR_STATE (-3)     /* call sequence violated (e.g. solve before mesh_set) */
#define DDPS_ERR_NOCONV (-4)    /* iteration cap reached (the reference would spin forever) */
#define DDPS_ERR_BREAKDOWN (-5) /* PCG breakdown: operator not SPD (u or v went negative) */
#define DDPS_ERR_COMM (-6)      /* NCCL failure */

/* ---- which Dirichlet value set (src:640 V; src:532-544 u and v; src:1001-1013 semlin) ----- */
enum ddps_field { DDPS_FIELD_V = 0, DDPS_FIELD_U = 1, DDPS_FIELD_W = 2 /* the reference's v */, DDPS_FIELD_SCALAR = 3 };

/* ---- sampled coefficient slots (host evaluates the closures once) ------ */
enum ddps_coeff {
  DDPS_COEFF_LAMBDA2 = 0, /* [nme]    lambda(centroid)^2                     src:250-254 */
  DDPS_COEFF_MU_P = 1,    /* [nme]    mu_p(centroid)                         src:463-467 */
  DDPS_COEFF_MU_N = 2,    /* [nme]    mu_n(centroid)                         src:463-467 */
  DDPS_COEFF_C_MID = 3,   /* [3,nme]  C at midpoints m12,m23,m31 (col-major) src:303-308,341-343 */
  DDPS_COEFF_A11 = 4,     /* [nme]    A[1](geom tag)                         src:901 */
  DDPS_COEFF_A22 = 5,     /* [nme]    A[4](geom tag)                         src:902 */
  DDPS_COEFF_H_MID = 6,   /* [3,nme]  u-independent part of f at midpoints   src:940-957 */
  DDPS_COEFF_NEUMANN = 7, /* [nq]     accumulated Neumann edge loads         src:986-996 */
  DDPS_COEFF_POLY_MID0 = 16,  /* +k: [3,nme] polynomial coefficient c_k at midpoints, k=0..4 */
  DDPS_COEFF_POLY_NODE0 = 32  /* +k: [nq]    polynomial coefficient c_k at nodes,     k=0..4 */
};

/* ---- registered u-dependent functors for solve_semlin_poisson's f,g (no CPU fallback) ----- */
enum ddps_functor_kind {
  DDPS_FUNCTOR_NONE = 0,
  DDPS_FUNCTOR_SINH = 1, /* f = H_MID + alpha*sinh(beta*u), g = alpha*beta*cosh(beta*u); params {alpha,beta} */
  DDPS_FUNCTOR_POLY = 2  /* f = sum_k c_k u^k, g = sum_k k c_k u^(k-1); params {degree} */
};

/* ---- registered SPATIAL functors: coefficient closures the library samples ON THE DEVICE from the resident mesh, at
 * the points where the reference evaluates them (centroids src:250-254,463-467; edge midpoints src:303-308,341-343;
 * nodes), bitwise equal to host sampling of the same expression.  Anything else is sampled by the host shim and uploaded
 * with ddps_coeff_set. */
enum ddps_spatial_kind {
  DDPS_SPATIAL_CONST = 1,  /* c                                  params {c} */
  DDPS_SPATIAL_AFFINE = 2, /* (a + bx*x) + by*y                   params {a, bx, by} */
  DDPS_SPATIAL_STEP_X = 3, /* x > x0 ? right : left               params {x0, left, right}   (pn doping, src:341) */
  DDPS_SPATIAL_STEP_Y = 4, /* y > y0 ? above : below              params {y0, below, above} */
  DDPS_SPATIAL_BY_TAG = 5  /* value of the triangle's geometric tag (Mesh.elements row 5, src:17); params
                              {default, tag_1, value_1, ..., tag_k, value_k}, k <= 7; per-triangle slots only */
};

enum ddps_rec_mode { DDPS_REC_ZERO = 0 /* :zero src:520 */, DDPS_REC_SHOCKLEY = 1 /* :shockley src:516 */ };

/* ---- matrices / vectors retrievable for parity tests -------------------------------------- */
enum ddps_matrix {
  DDPS_MAT_BASE = 0, /* eliminated constant stiffness: MV (src:284) or semlin M (src:1028) */
  DDPS_MAT_OP = 1    /* last assembled operator M - J0 handed to the solver (src:394,588,1063) */
};
enum ddps_vector {
  DDPS_VEC_RHS = 0,   /* last assembled right-hand side b (src:361-363, 562-564, 1031-1033) */
  DDPS_VEC_RESID = 1, /* last Newton residual M*V - b (src:368,1040) */
  DDPS_VEC_V = 2, DDPS_VEC_U = 3, DDPS_VEC_W = 4, /* current fields */
  DDPS_VEC_LIFT = 5   /* K[:,D]*gD of the base stiffness (src:362, rows not in D) */
};

typedef struct ddps_solver_opts {
  double lin_rtol;      /* PCG stop for the linear continuity solves: ||r||_2 <= lin_rtol*||b||_2   (default 1e-13) */
  double newton_eta;    /* PCG stop for Newton corrections: ||r||_inf <= 0.1*tol (so a linear problem converges in one
                           Newton step like the reference's direct solve) OR ||r||_2 <= newton_eta*||F||_2 (default 1e-15) */
  int64_t lin_max_it;   /* PCG iteration cap per solve (default 200000) */
  int32_t max_newton;   /* Newton cap per solve, 0 = 1000 (the reference has none, src:368,1040) */
  int32_t max_gummel;   /* Gummel cap, 0 = 1000 (the reference has none, src:649) */
  int32_t check_every;  /* PCG iterations between host polls of the device-side done flag (default 32) */
  int32_t warm_start;   /* 1: linear solves start PCG from what the previous Gummel iteration found -- the continuity solves from
                           the previous iterate, the k-th Newton correction of a Vsolve from the k-th correction of the previous
                           Vsolve (same solutions, same Newton iterates and counts: only fewer CG iterations) (default 1) */
  int32_t verbose;      /* 1: print "iteration k, tolerance: c" lines like src:655,1093 */
  int32_t lin_cap_ok;   /* 1: PCG / Newton solves that hit their caps are accepted instead of failing (profiling runs only) */
  int32_t no_alt_traversal; /* 1: disable the alternating traversal direction of the PCG kernels (A/B timing) */
  int32_t warm_start_v; /* OPT-IN, changes the iteration history: Vsolve starts Newton from the previous Gummel iterate
                           instead of V=0 (src:311) -- SURVEY section 8f rank 4.  Default 0 = reference behaviour */
  int32_t no_sym_scaling; /* 1: classic Jacobi-PCG with explicit D^-1 reads instead of CG on D^-1/2 A D^-1/2 (same iterates) */
  int32_t renumber;     /* set BEFORE ddps_mesh_set.  1: renumber the nodes internally along a Morton (Z-order) curve so that
                           mesh neighbours are index neighbours (gmsh numbering is insertion order); the ABI keeps the
                           caller's numbering for every input and output.  Default 0 */
  int32_t precond;      /* preconditioner of the hand-written CG.  DDPS_PRECOND_JACOBI (0, default):
                           the Jacobi-PCG the north star names.  DDPS_PRECOND_AMG (1): smoothed-aggregation multigrid
                           V(1,1) cycle (frozen prolongators from the base stiffness, Galerkin operators of every assembled
                           M - J0 recomputed on the device); same stop rules, same solutions to solver tolerance.
                           With a communicator the hierarchy is distributed: rank-local aggregates, halo exchanges of the
                           cycle vectors over peer memory on the large levels, a replicated tail below ~200k rows */
  int32_t amg_no_lag;   /* multigrid path: 1 = recompute the coarse (Galerkin) operators for every assembled operator.  Default 0:
                           they are reused while the operator of the same kind (first / later Newton system, u-, v-continuity
                           system) stays within 1 % (entries and row sums) of the one they were computed from -- the cycle is
                           only a preconditioner, the solutions are the same to solver tolerance */
  int32_t newton_damping; /* OPT-IN, default 0 = the reference's undamped update (src:394-395,
                           1063-1064).  k > 0: the Newton step of Vsolve / solve_semlin_poisson is halved, at most k times,
                           until the residual max-norm ||M*V - b(V)||_inf (the quantity the loops test, src:368,1040) decreases.
                           Changes the iteration history only when a full step would increase the residual */
  int32_t reserved0;
  double lin_etol;      /* error-based stop rule of every linear solve inside the Newton / Gummel loops, on top of the residual
                           rules above: the A-norm of the error (multigrid path: r.Br; Jacobi path: Hestenes-Stiefel estimate
                           over the last 32 updates) must fall below eps * sqrt(E), E = the energy of the solution (x0.b of a
                           warm start / the largest Newton correction of the running loop), eps = max(1e-13, min(lin_etol,
                           0.01*tol)).  Default 1e-10; < 0 switches the rule off (residual rules only).  The reference's
                           direct solves are exact to rounding; a residual-only rule leaves an error ~ tol/h^2 behind, which
                           breaks the 1e-8 field parity on large meshes */
} ddps_solver_opts;

#define DDPS_PRECOND_JACOBI 0
#define DDPS_PRECOND_AMG 1

typedef struct ddps_stats {
  int32_t outer_iters;     /* Gummel (ddpe) or Newton (semlin) iterations done */
  int32_t newton_total;    /* Newton iterations summed over all Vsolve calls */
  int64_t pcg_total;       /* PCG iterations summed over all solves */
  int64_t spmv_total;      /* SpMV launches (PCG + residuals) */
  int64_t kernel_launches; /* all kernel launches of this library during the call */
  double control;          /* last Gummel update norm (src:653) or last Newton residual (src:1092) */
  double t_setup_ms;       /* host wall: uploads + pattern build inside this call (0 when resident) */
  double t_solve_ms;       /* device-resident loop, CUDA events on the library stream */
  double t_asm_ms;         /* summed CUDA-event time of assembly kernels */
  double t_pcg_ms;         /* summed CUDA-event time of PCG */
  double t_wall_ms;        /* host wall clock around the same region as t_solve_ms */
  int32_t status;          /* DDPS_OK / DDPS_ERR_NOCONV / DDPS_ERR_BREAKDOWN */
  int32_t n_negative_area; /* elements with signed area < 0 (Q2: mass uses signed ar, src:375) */
  int32_t amg_levels;      /* multigrid levels in use (0: Jacobi-PCG) */
  int32_t amg_reuses;      /* solves that reused lagged coarse operators (opts.amg_no_lag = 0) */
  double t_amg_setup_ms;   /* host wall: one-time hierarchy setup (aggregation, prolongators, coarse patterns, uploads) */
  int32_t newton_halvings; /* step halvings taken by opts.newton_damping */
  int32_t reserved[3];
} ddps_stats;

/* ---- lifecycle --------------------------------------------------------------------------- */
int ddps_version(void);
const char* ddps_last_error(const ddps_ctx* ctx);
int ddps_ctx_create(ddps_ctx** out, int device);
int ddps_ctx_destroy(ddps_ctx* ctx);
void ddps_default_opts(ddps_solver_opts* opts);
int ddps_set_opts(ddps_ctx* ctx, const ddps_solver_opts* opts);

/* ---- multi-GPU (one process per GPU; node-partitioned by coordinate bisection) ------------ */
/* The reference has no communication at all; these are new.  id is an ncclUniqueId (128 bytes)
 * produced on rank 0 and broadcast by the host's own plumbing (torch.distributed / MPI / files).
 * An id that starts with "LOCAL:" selects the in-process backend instead: the ranks are THREADS of one process sharing one
 * device (each with its own ctx), exchanges are device-to-device copies -- the whole distributed path (partition, ghost
 * layers, distributed multigrid) runs on a single GPU; this is what the single-GPU test-suite drives. */
int ddps_comm_unique_id(char id128[128]);
int ddps_comm_init(ddps_ctx* ctx, const char id128[128], int rank, int nranks);

/* ---- mesh: replaces the per-call gathers of mesh.nodes/mesh.elements (src:235-247, 443-456, 909-921)
 * nodes    f64 [2,nq]  column-major (x,y per node)        = Mesh.nodes    (src:11)
 * elements i64 [5,nme] column-major, 1-based node ids     = Mesh.elements (src:17)
 * Builds the fixed sparsity pattern + row->incident-triangle lists once (replaces every sparse()
 * call, src:275,284,361,391,...).  With a communicator every rank passes the SAME global mesh. */
int ddps_mesh_set(ddps_ctx* ctx, const double* nodes, int64_t nq, const int64_t* elements, int64_t nme);

/* Per-rank mesh ingest for multi-GPU runs: every rank passes ONLY its part, nobody holds the global mesh (at 268M triangles
 * the global arrays are 13 GB per process).  Any node partition is accepted.
 *   nodes      f64 [2,n_loc]   the n_own nodes this rank owns first, in increasing global id, then its ghost nodes (= every
 *                              other node of a triangle that touches an owned node), sorted by (owner rank, global id)
 *   node_glob  i64 [n_loc]     1-based global id of every local node
 *   ghost_owner i32 [n_loc-n_own]  rank that owns each ghost node
 *   elements   i64 [5,nme_loc] EVERY triangle that touches an owned node, LOCAL 1-based node ids, increasing global
 *                              triangle id (= the reference's summation order per matrix entry)
 *   elem_glob  i64 [nme_loc]   1-based global triangle ids, or NULL (then coefficient slots can only be filled by
 *                              ddps_coeff_functor, not from global arrays)
 * Dirichlet sets keep GLOBAL node ids (every rank passes the full, small boundary list).  Results: ddps_ddpe_fetch_owned. */
int ddps_mesh_set_local(ddps_ctx* ctx, const double* nodes, int64_t n_own, int64_t n_loc, const int64_t* node_glob,
                        const int32_t* ghost_owner, const int64_t* elements, int64_t nme_loc, const int64_t* elem_glob,
                        int64_t nq_global, int64_t nme_global);
int ddps_ddpe_fetch_owned(ddps_ctx* ctx, double* V, double* u, double* v, int64_t cap);

/* Host-only partition query (runs without a GPU context's kernels; used by the CPU tests):
 * owner[nq] <- rank owning each node under recursive coordinate bisection into nparts parts. */
int ddps_partition_nodes(const double* nodes, int64_t nq, int nparts, int32_t* owner);

/* Host-only: the local mesh + halo plan rank `rank` of `nparts` derives from the global mesh (the same
 * code ddps_mesh_set runs).  counts = {n_own, n_ghost, nme_local, n_neighbours}; loc2glob[n_own+n_ghost]
 * (0-based global ids: owned first, then ghosts grouped by owner); peers/send_counts/recv_counts
 * [n_neighbours]; send_glob = global ids this rank sends, grouped by peer in peers[] order.  Any
 * output pointer may be NULL.  Capacity nq (loc2glob, send_glob) / nparts (the others) suffices. */
int ddps_partition_plan(const double* nodes, int64_t nq, const int64_t* elements, int64_t nme, int nparts, int rank,
                        int64_t counts[4], int64_t* loc2glob, int32_t* peers, int64_t* send_counts, int64_t* recv_counts,
                        int64_t* send_glob);

/* ---- boundary + coefficient uploads (host shim output of aquire_boundary, src:162-218) ---- */
int ddps_dirichlet_set(ddps_ctx* ctx, int field, const int64_t* idx1, const double* val, int64_t nd);
int ddps_coeff_set(ddps_ctx* ctx, int slot, const double* data, int64_t len);
int ddps_functor_set(ddps_ctx* ctx, int kind, const double* params, int nparams);
/* Fill a coefficient slot by evaluating a registered spatial functor on the device (replaces the reference's per-call
 * `lambda.(cx,cy)`, `map(mu, ...)`, `C.(halfa_x, halfa_y)` broadcasts, src:254,467,341-343): no host sampling, no upload.
 * DDPS_COEFF_LAMBDA2 stores the SQUARE of the functor's value (src:254). */
int ddps_coeff_functor(ddps_ctx* ctx, int slot, int kind, const double* params, int nparams);

/* ---- solve_ddpe (src:630-658) -------------------------------------------------------------
 * begin: MV_assemble once (src:643), zero state (src:635-637).
 * step : nsteps applications of the Gummel map G (src:596-601, 649-656); control[k] = update norm
 *        of step k (src:653); stops early when control <= tol if stop_at_tol != 0.
 * fetch: copy the current V,u,v (length nq each) to the host.
 * solve: begin + step until control <= tol + fetch  == solve_ddpe(...) -> (V,u,v). */
int ddps_ddpe_begin(ddps_ctx* ctx, int rec_mode, double delta, double tau_p, double tau_n, double tol);
int ddps_ddpe_step(ddps_ctx* ctx, int nsteps, int stop_at_tol, double* control, int* steps_done);
int ddps_ddpe_fetch(ddps_ctx* ctx, double* V, double* u, double* v);
int ddps_solve_ddpe(ddps_ctx* ctx, int rec_mode, double delta, double tau_p, double tau_n, double tol,
                    double* V, double* u, double* v, ddps_stats* stats);

/* ---- solve_semlin_poisson (src:895-1097): Newton on M*u = b(u), Jacobian M - J0(g) -------- */
int ddps_solve_semlin(ddps_ctx* ctx, double tol, double* V, ddps_stats* stats);

/* ---- iteration history of the last solve ---------------------------------------------------
 * kind 0: Gummel control per outer iteration (src:653) / Newton residual per iteration (src:1092)
 * kind 1: Newton iterations per Vsolve call; kind 2: PCG iterations per linear solve. */
int ddps_get_history(ddps_ctx* ctx, int kind, double* out, int64_t cap, int64_t* n);
int ddps_get_stats(ddps_ctx* ctx, ddps_stats* stats);

/* ---- parity / debug exports (single-GPU contexts) ------------------------------------------ */
int ddps_get_csr_size(ddps_ctx* ctx, int64_t* nrows, int64_t* nnz);
/* CSR of the fixed pattern, 0-based, columns sorted within each row, explicit zeros kept (Q5). */
int ddps_get_csr(ddps_ctx* ctx, int which, int64_t* rowptr, int64_t* col, double* val);
int ddps_get_vector(ddps_ctx* ctx, int which, double* out, int64_t len);
/* The coefficient array a slot holds on the device (uploaded or device-sampled), in the layout ddps_coeff_set takes;
 * *len receives its length (out may be NULL to query it). */
int ddps_get_coeff(ddps_ctx* ctx, int slot, double* out, int64_t cap, int64_t* len);
/* Assemble individual systems from host-supplied fields (each length nq), results retrievable with
 * ddps_get_csr(DDPS_MAT_OP)/ddps_get_vector(DDPS_VEC_RHS):
 *   base   : MV (src:229-287) from LAMBDA2, or semlin M (src:924-1028) from A11/A22 (use_semlin!=0)
 *   newton : A = base - J0(V), b(V), F = base*V - b   (Vsolve src:368-420 / semlin src:1040-1089)
 *   uv     : A,b of uvsolve (src:433-588); which = DDPS_FIELD_U (V,tau_p,tau_n,mu_p,u0,v0) or
 *            DDPS_FIELD_W (-V,tau_n,tau_p,mu_n,v0,u0) exactly as G calls it (src:598-599). */
int ddps_debug_assemble_base(ddps_ctx* ctx, int use_semlin);
int ddps_debug_assemble_newton(ddps_ctx* ctx, int use_semlin, double delta, const double* V, const double* u,
                               const double* v, double* resid_inf);
int ddps_debug_assemble_uv(ddps_ctx* ctx, int which, int rec_mode, double tau_p, double tau_n, const double* V,
                           const double* u0, const double* v0);
/* y = A*x with the chosen matrix (K8). */
int ddps_debug_spmv(ddps_ctx* ctx, int which, const double* x, double* y);
/* Solve OP*x = rhs with Jacobi-PCG (K9); x holds the initial guess on entry. */
int ddps_debug_pcg(ddps_ctx* ctx, const double* rhs, double* x, double rtol, int64_t max_it, int64_t* iters,
                   double* relres);

/* ---- in-library kernel timing (CUDA events on the library's own stream) --------------------
 * kernel 0: SpMV on DDPS_MAT_OP; 1: continuity-system assembly (uvsolve A+b, src:433-588);
 * 2: one full PCG iteration; 3: Newton-system assembly (Vsolve, src:368-420); 4: base stiffness;
 * 5: the PCG's SpMV with the fused p.Ap dot; 6: one CG iteration preconditioned with the multigrid V-cycle
 * (opts.precond = DDPS_PRECOND_AMG); 7: the per-operator multigrid numeric phase (Galerkin products on all levels +
 * coarsest inverse); 8: one V-cycle alone; single kernels of the multigrid iteration on the finest level: 9 residual SpMV
 * t = b - A x, 10 smoothing-sweep SpMV fused with r.z, 11 the CG update kernel (x, r, first sweep), 12 restriction to level 1,
 * 13 prolongation from level 1 (kernel 6 reports the summed algorithmic bytes of the whole iteration); 14 / 15 the nodal prep pass
 * of the Newton-system / continuity-system assembly alone (kernels 1 and 3 time the prep pass and the assembly kernel together).
 * Runs `reps` launches after `warm` untimed ones and returns the mean milliseconds per launch and
 * the algorithmic bytes of one launch (definition in DESIGN.md). */
int ddps_time_kernel(ddps_ctx* ctx, int kernel, int warm, int reps, double* ms_per_launch, double* algo_bytes);

/* ---- multigrid preconditioner (replaces `\` at src:394,588,1063 together with the CG) ----
 * Host-only hierarchy setup, callable WITHOUT a GPU (the CPU tests check it against scipy): from a CSR matrix
 * (0-based, sorted columns) and a per-row `fixed` flag (Dirichlet rows) build the smoothed-aggregation hierarchy
 * the library uses internally.  level_sizes: {n, nnz, n_coarse (0 on the coarsest level), nnz(P)}.
 * level_get: CSR of the level's (base) operator, CSR of the prolongator P [n x n_coarse], aggregate id per row
 * (-1: none); any pointer may be NULL. */
typedef struct ddps_amg_host ddps_amg_host;
int ddps_amg_host_create(ddps_amg_host** out, int64_t n, const int64_t* rowptr, const int64_t* col, const double* val,
                         const uint8_t* fixed, int nthreads, int coarse_max);
int ddps_amg_host_nlevels(const ddps_amg_host* h);
int ddps_amg_host_level_sizes(const ddps_amg_host* h, int level, int64_t sizes[4]);
int ddps_amg_host_level_get(const ddps_amg_host* h, int level, int64_t* rowptr, int64_t* col, double* val, int64_t* Pptr,
                            int64_t* Pcol, double* Pval, int64_t* agg);
void ddps_amg_host_free(ddps_amg_host* h);
/* Device-side parity/debug exports (single-GPU contexts, opts.precond = DDPS_PRECOND_AMG, after a base assembly):
 * levels : number of levels and rows per level (sizes capacity 16);
 * numeric: run the per-operator numeric phase on DDPS_MAT_OP / DDPS_MAT_BASE;
 * level  : the hierarchy as it lives on the device: sizes = {n, nnz, n_coarse (0 on the coarsest level), nnz(P)}, CSR pattern
 *          of the level (sorted columns) and its prolongator P [n x n_coarse] (the library stores P in fp32); any array
 *          pointer may be NULL;
 * values : CSR-ordered values + inverse diagonal of level >= 1 after the last numeric phase;
 * apply  : z = one V(1,1) cycle applied to r (length nq each, caller's numbering). */
int ddps_amg_levels(ddps_ctx* ctx, int64_t* nlevels, int64_t* sizes);
int ddps_amg_numeric(ddps_ctx* ctx, int which);
int ddps_amg_get_level(ddps_ctx* ctx, int level, int64_t sizes[4], int64_t* rowptr, int64_t* col, int64_t* Pptr, int64_t* Pcol,
                       double* Pval);
int ddps_amg_get_values(ddps_ctx* ctx, int level, double* val, double* dinv);
int ddps_amg_apply(ddps_ctx* ctx, const double* r, double* z);

#ifdef __cplusplus
}
#endif
#endif /* DDPS_H */
