/* dzahui_b200.h — C ABI of the B200-native (sm_100a) 1D-FEM hot path of Dzahui.
 *
 * The reference (Arthur-phys/Dzahui, pure Rust) has no FFI: its seam is the trait object
 * `Box<dyn DiffEquationSolver>` built in DzahuiWindow::run (simulation/dzahui_window.rs:738-800) and called
 * once per frame (`solver.solve(self.time_step)`, :976).  This header is what a Rust shim implementing that
 * trait binds (INTEGRATION.md shows the `extern "C"` block and the `impl DiffEquationSolver`).  Each entry
 * point names the reference interface it replaces.  Paths are relative to /root/reference/src/lib.
 *
 * Conventions: plain pointers and sizes only; every function returns a dzb_status (0 = OK) and records a
 * message retrievable with dzb_last_error(); handles are not thread-safe (the reference calls everything from
 * the winit event-loop thread).  All compute runs on the current CUDA device; there is NO CPU fallback: without
 * a usable GPU every compute entry point returns DZB_CUDA.
 */
#ifndef DZAHUI_B200_H
#define DZAHUI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors the variants of `Error` (error.rs:32-55) that the hot path can raise. */
typedef enum dzb_status {
  DZB_OK = 0,
  DZB_WRONG_DIMS = 1,     /* Error::WrongDims        (time_dependent.rs:66-68, matrix_solver/mod.rs:19-21, utils.rs:43-45) */
  DZB_INTEGRATION = 2,    /* Error::Integration      (quadrature/gauss_legendre.rs:5923-5925) */
  DZB_PIECEWISE_DIMS = 3, /* Error::PieceWiseDims    (piecewise_polynomials_1degree.rs:112-114) */
  DZB_MESH_PARSE = 4,     /* Error::MeshParse / ParseFloat / Infallible while reading the .obj (mesh_builder.rs:58-84,140-190,217-227) */
  DZB_IO = 5,             /* Error::Io               (mesh_builder.rs:141,211) */
  DZB_CUDA = 6,           /* no reference equivalent: CUDA runtime failure / no device */
  DZB_INVALID = 7         /* no reference equivalent: NULL handle, bad option (the Rust builders `panic!` instead) */
} dzb_status;

typedef struct dzb_mesh dzb_mesh;         /* mesh::Mesh (mesh/mod.rs:30-37), 1D part, device resident */
typedef struct dzb_solver dzb_solver;     /* Box<dyn DiffEquationSolver> (solvers/solver_trait.rs:12-23) */
typedef struct dzb_mesh2d dzb_mesh2d;     /* mesh::Mesh of a planar triangle mesh (viewer only in the reference) */
typedef struct dzb_ensemble dzb_ensemble; /* many independent DiffussionSolverTimeDependent bars stepped together */

/* How the element integrals are evaluated on the device. */
typedef enum dzb_assembly_mode {
  DZB_ASSEMBLY_AUTO = 0,     /* static solvers: FAITHFUL (nodal parity on long bars needs the reference's bands bit for bit); time dependent: FAITHFUL while nodes*gauss_step <= 2^24, FAST beyond */
  DZB_ASSEMBLY_FAITHFUL = 1, /* the reference's quadrature loop, same operation order, no FMA: bands bit-identical to the reference arithmetic */
  DZB_ASSEMBLY_FAST = 2      /* closed form over prefix moments of the same rule (incl. the dropped node and the split at the kink): O(log G) per node, differs by rounding only */
} dzb_assembly_mode;

/* How the tridiagonal systems are solved on the device. */
typedef enum dzb_solve_mode {
  DZB_SOLVE_AUTO = 0,     /* FAITHFUL for small problems (<= 2^16 unknowns in total), PARALLEL beyond */
  DZB_SOLVE_FAITHFUL = 1, /* matrix_solver::solve_by_thomas (mod.rs:17-48) operation for operation, one thread per bar */
  DZB_SOLVE_PARALLEL = 2  /* ensembles: pre-factored Thomas, per-lane chunk sweeps + warp-shuffle scans of the affine recurrences; long bars: recursive Thomas/PCR-style chunk reduction (registers -> shared memory -> grid), used when the matrix is an M-matrix or strictly dominant (checked on the device after every assembly), the serial recurrence otherwise */
} dzb_solve_mode;

typedef struct dzb_options {
  int assembly_mode; /* dzb_assembly_mode */
  int solve_mode;    /* dzb_solve_mode */
  int refine_steps;  /* static solvers, PARALLEL mode: iterative-refinement sweeps with a double-double residual (-1 = auto = 0: the unrefined solve is within 1e-11 of the exact solution of the system at 10^7 rows; one sweep reaches 1e-18) */
  int stored_operators; /* ensembles, PARALLEL mode: 0 = operator rebuilt on chip (k_td_step_lean, 40 B per DOF-step), 1 = pre-scaled operator arrays (k_td_step_fast, 56 B) */
} dzb_options;

/* ---- library / device ------------------------------------------------------------------------------------ */
const char* dzb_last_error(void);
int dzb_version(void);
int dzb_device_count(int* count);
int dzb_set_device(int device);            /* one process per GPU: call with LOCAL_RANK before anything else */
int dzb_set_stream(void* cuda_stream);     /* cudaStream_t all subsequent work is enqueued on (default: stream 0) */
int dzb_synchronize(void);
/* page-locked host memory for the buffers handed to dzb_*_solve / dzb_*_state / dzb_*_read: copies into it run at PCIe speed;
 * any other host pointer works too, only slower (the driver stages pageable memory) */
int dzb_host_alloc(void** ptr, size_t bytes);
int dzb_host_free(void* ptr);
/* number of kernels this library has launched since the last reset (bench.py's gpu_launches) */
int dzb_launch_count(unsigned long long* count, int reset);

/* ---- mesh ingestion: MeshBuilder::build_mesh_1d (mesh/mesh_builder.rs:204-328, with :58-84 and :140-190) ---- */
/* Parses the `v x y z` lines, picks the varying axis, insertion-sorts ascending, and leaves on the device:
 * node coordinates f64[n], GL vertices f64[12n], triangle indices u32[6n-6] and the element-to-node table
 * u32[2(n-1)] (element e = (e, e+1): implicit in the reference), all bit-identical to the reference's indexing. */
int dzb_mesh_ingest_obj_1d(const char* path, int has_height_multiplier, double height_multiplier, dzb_mesh** out);
/* `mesh: Vec<f64>` handed straight to XSolver::new (time_independent.rs:57): sorted ascending, n >= 3. */
int dzb_mesh_from_nodes(const double* x, size_t n, dzb_mesh** out);
/* Re-points an existing from_nodes mesh at new coordinates (same n) without reallocating: a moving / re-meshed bar. */
int dzb_mesh_set_nodes(dzb_mesh* m, const double* x, size_t n);
int dzb_mesh_node_count(const dzb_mesh* m, size_t* n);
int dzb_mesh_nodes(const dzb_mesh* m, double* out, size_t n);         /* Mesh::filter_for_solving_1d (mesh/mod.rs:54-68) */
int dzb_mesh_vertices(const dzb_mesh* m, double* out, size_t len);    /* Mesh.vertices, len = 12n  (0 for from_nodes meshes) */
int dzb_mesh_indices(const dzb_mesh* m, uint32_t* out, size_t len);   /* Mesh.indices,  len = 6n-6 */
int dzb_mesh_elements(const dzb_mesh* m, uint32_t* out, size_t len);  /* element-to-node, len = 2(n-1) */
int dzb_mesh_info(const dzb_mesh* m, double* max_length, double* prom_width, float translation[3]); /* :270,:277,:310-318 */
int dzb_mesh_device_nodes(const dzb_mesh* m, const double** device_ptr);
/* Mesh::update_gradient_1d (mesh/mod.rs:73-90) fused with Drawable::get_vertices' f64->f32 cast (:108-113):
 * colours the device-resident vertex buffer from |solution| and returns the f32[12n] buffer a renderer uploads. */
int dzb_mesh_update_gradient_1d(dzb_mesh* m, const double* device_solution, size_t n, float* out_vertices_f32, size_t len);
void dzb_mesh_destroy(dzb_mesh* m);

/* ---- 2D ingestion: MeshBuilder::build_mesh_2d (mesh/mesh_builder.rs:342-500, obj_face_checker :89-129, merge_sort :623-680) ---- */
/* Parses `v` and `f a/b/c` lines, drops the constant axis, and leaves on the device the GL vertices f64[6V] ([a,b,0,0,0,1]),
 * the 0-based triangle indices u32[3T] and the ascending, duplicate-free list of boundary vertices (end points of the
 * edges that belong to exactly one triangle), found with a device hash table instead of the reference's HashMap + sort. */
int dzb_mesh_ingest_obj_2d(const char* path, dzb_mesh2d** out);
int dzb_mesh2d_counts(const dzb_mesh2d* m, size_t* n_vertices, size_t* n_triangles, size_t* n_boundary);
int dzb_mesh2d_vertices(const dzb_mesh2d* m, double* out, size_t len);           /* Mesh.vertices, len = 6V */
int dzb_mesh2d_indices(const dzb_mesh2d* m, uint32_t* out, size_t len);          /* Mesh.indices,  len = 3T */
int dzb_mesh2d_boundary_indices(const dzb_mesh2d* m, uint32_t* out, size_t len); /* Mesh.boundary_indices */
int dzb_mesh2d_info(const dzb_mesh2d* m, double* max_length, float translation[3]);
void dzb_mesh2d_destroy(dzb_mesh2d* m);
/* the boundary detection alone on a device-resident triangle list; d_out needs room for n_vertices entries */
int dzb_boundary_vertices_device(const uint32_t* d_triangles, size_t n_triangles, size_t n_vertices, uint32_t* d_out, size_t* n_out);

/* ---- quadrature: gauss_legendre::quad_pair (quadrature/gauss_legendre.rs:5908-5927) ------------------------- */
int dzb_quad_pair(size_t n, size_t k, double* theta, double* weight);
/* the (cos(theta_j), w_j), j = 1..gauss_step-1 that every integration loop consumes (time_independent.rs:132-135) */
int dzb_quad_table(size_t gauss_step, double* x, double* w);

/* ---- solvers: XSolver::new(&params, mesh, gauss_step) ------------------------------------------------------ */
/* StokesSolver1D::new (stokes_solver/dim1.rs:77-94).  `force` is sampled on the host, during this call only, at
 * exactly the reference's points and in its order (interior rows i = 1..n-2 with j ascending, then row 0). */
int dzb_stokes1d_create(const dzb_mesh* m, double rho, double hydrostatic_pressure,
                        double (*force)(double x, void* ctx), void* ctx, size_t gauss_step,
                        const dzb_options* opt, dzb_solver** out);
/* DiffussionSolverTimeIndependent::new (diffusion_solver/time_independent.rs:57-71) */
int dzb_diffusion_ti_create(const dzb_mesh* m, double mu, double b, double bc_left, double bc_right,
                            size_t gauss_step, const dzb_options* opt, dzb_solver** out);
/* DiffussionSolverTimeDependent::new (diffusion_solver/time_dependent.rs:62-95): DZB_WRONG_DIMS unless n_ic == n-2 */
int dzb_diffusion_td_create(const dzb_mesh* m, double mu, double b, double bc_left, double bc_right,
                            const double* initial_conditions, size_t n_ic, size_t gauss_step,
                            const dzb_options* opt, dzb_solver** out);

/* XSolver::new again, into the buffers the handle already owns (same node count, same params): re-assembles K (or M, S;
 * the time-dependent state is kept) for the mesh's current coordinates.  This is the "assemble" half of one
 * assemble + solve step without allocator traffic; the Stokes solver (host closure) is not rebuildable. */
int dzb_solver_rebuild(dzb_solver* s, const dzb_mesh* m);

/* DiffEquationSolver::solve(&mut self, time_step) -> Result<Vec<f64>, Error> (solver_trait.rs:23).
 * Synchronous: `out_host[0..n)` holds the nodal values on return.  Static solvers ignore time_step; the
 * time-dependent solver advances its state by exactly one step of that size (it may change between calls). */
int dzb_solver_solve(dzb_solver* s, double time_step, double* out_host, size_t n);
/* Same step, result left on the device (valid until the next call on this handle): for a renderer that only
 * reads back when it draws. */
int dzb_solver_solve_device(dzb_solver* s, double time_step, const double** device_ptr);
int dzb_solver_node_count(const dzb_solver* s, size_t* n);
/* which kernels the last dzb_solver_solve[_device] of this handle ran through (what bench.py prints next to its numbers) */
int dzb_solver_path(const dzb_solver* s, char* buf, size_t cap);
/* `stiffness_matrix` / `mass_matrix` / `b_vector` (pub(crate) fields the reference tests index as [[i,j]]):
 * which = 0 -> stiffness (K or S), 1 -> mass (time-dependent only).  lo[i] = A[i][i-1], up[i] = A[i][i+1];
 * rhs may be NULL (and is ignored for the time-dependent solver). */
int dzb_solver_bands(const dzb_solver* s, int which, double* lo, double* di, double* up, double* rhs, size_t n);
int dzb_solver_state(const dzb_solver* s, double* out, size_t n);  /* `state` (time_dependent.rs:55) */
void dzb_solver_destroy(dzb_solver* s);

/* ---- one bar split across GPUs (SPIKE): every rank owns a contiguous block of rows ---------------------------- */
/* DiffussionSolverTimeIndependent::new (time_independent.rs:57-71) for the rows of one block.  `m` holds the block's own
 * nodes plus one halo node on each side that has a neighbour (has_left / has_right): assembly of an interior row needs
 * x_{i-1}, x_i, x_{i+1} only.  The first block carries the left Dirichlet row, the last block the right one. */
int dzb_diffusion_ti_create_block(const dzb_mesh* m, int has_left, int has_right, double mu, double b, double bc_left,
                                  double bc_right, size_t gauss_step, const dzb_options* opt, dzb_solver** out);
/* Eliminates everything inside the block; top8 = its two interface rows (a, c, s, f) x 2 in row-excess form
 * (diagonal = s - a - c).  which_rhs: 0 = b_vector, 1 = the residual left by dzb_solver_block_residual. */
int dzb_solver_block_reduce(dzb_solver* s, int which_rhs, double top8[8]);
/* After the ranks exchanged their rows (ONE all-gather of 8 doubles per rank) and solved the interface system:
 * back-substitute with this block's end unknowns; accumulate != 0 adds to the stored solution (refinement). */
int dzb_solver_block_backsolve(dzb_solver* s, int which_rhs, double us, double ue, int accumulate);
int dzb_solver_block_edges(const dzb_solver* s, double edges[2]);              /* first / last value of the block's solution */
int dzb_solver_block_residual(dzb_solver* s, double u_left, double u_right);   /* r = b - K u (double-double) with halo values */
int dzb_solver_block_read(const dzb_solver* s, double* out_host, size_t n);
/* rows[8 r ..] = block r's top8; ends[2 r], ends[2 r + 1] = the values to hand to block r's back-substitution */
int dzb_interface_solve(const double* rows, size_t n_blocks, double* ends);
/* The same steps with every operand in DEVICE memory and no host synchronisation, so that reduce -> ncclAllGather (on the
 * stream given to dzb_set_stream) -> interface solve -> back-substitution queue up behind one another:
 * d_top8: 8 doubles; d_rows: 8 per block (all ranks, gathered); d_ends: 2 per block; d_work: 4 per block of scratch;
 * d_ends2: this block's pair inside d_ends; d_left / d_right: the neighbours' edge values (NULL where there is none). */
int dzb_solver_block_reduce_device(dzb_solver* s, int which_rhs, double* d_top8);
int dzb_interface_solve_device(const double* d_rows, size_t n_blocks, double* d_ends, double* d_work);
int dzb_solver_block_backsolve_device(dzb_solver* s, int which_rhs, const double* d_ends2, int accumulate);
int dzb_solver_block_edges_device(const dzb_solver* s, double* d_edges2);
int dzb_solver_block_residual_device(dzb_solver* s, const double* d_left, const double* d_right);
int dzb_solver_block_device_solution(const dzb_solver* s, const double** device_ptr);

/* ---- the split bar as ONE collective call per rank: the communicator lives in the library ---------------------- */
/* One process per GPU (at most 8 ranks: one NVSwitch box).  Rank 0 makes the id and ships its 128 bytes to the other
 * ranks by whatever the host has (MPI, a file, a socket); every rank then calls dzb_comm_init after dzb_set_device.
 * The library binds NCCL at run time (libnccl.so.2), so a Rust host needs no NCCL binding of its own.  Besides the NCCL
 * communicator, init maps a small mailbox of every rank into every other rank (cudaIpc*, NVLink peer access): where that
 * works (dzb_comm_info: peer_mailboxes = 1) the 8 doubles per rank of a solve travel as peer stores + a flag from INSIDE
 * the kernels and no collective is launched at all; otherwise they go through ncclAllGather on the library's stream. */
#define DZB_COMM_ID_BYTES 128
typedef struct dzb_comm dzb_comm;
int dzb_comm_unique_id(unsigned char id[DZB_COMM_ID_BYTES]);
int dzb_comm_init(const unsigned char id[DZB_COMM_ID_BYTES], int rank, int world, dzb_comm** out); /* collective; id may be NULL when world == 1 */
int dzb_comm_info(const dzb_comm* c, int* rank, int* world, int* peer_mailboxes);
int dzb_comm_all_gather(dzb_comm* c, const double* d_send, double* d_recv, size_t count_per_rank); /* device pointers, the library's stream */
void dzb_comm_destroy(dzb_comm* c);
/* XSolver::new again for a block (time_independent.rs:57-71), into the buffers it owns: `m` = the block's nodes + halo
 * nodes as at creation, current coordinates.  The "assemble" half of a step; a no-op but for a coordinate copy when the
 * block forms its rows inside the solve (DZB_ASSEMBLY_FAST). */
int dzb_solver_block_rebuild(dzb_solver* s, const dzb_mesh* m);
/* DiffEquationSolver::solve for the split bar (solver_trait.rs:23; ONE solve_by_thomas over the whole system in the
 * reference, matrix_solver/mod.rs:17-48).  Collective: every rank of `c` calls it with its block (rank r owns the r-th
 * block).  Block elimination, exchange of the interface rows, interface solve (redundantly on every rank) and
 * back-substitution are queued on the library's stream without any host synchronisation; with DZB_ASSEMBLY_FAST and peer
 * mailboxes the whole assemble + solve of a rank is ONE kernel launch.  refine_steps of the block's options add one
 * exchange of edge values + one more solve each.  *device_ptr (may be NULL) = the block's nodal values on the device. */
int dzb_solver_block_solve_device(dzb_solver* s, dzb_comm* c, const double** device_ptr);
/* Synchronises and reports what the collective solves since the last call left behind: DZB_CUDA when some rank's rows never
 * arrived (time-out instead of a hang), DZB_INVALID when new coordinates made the block's matrix unfit for the partitioned
 * elimination or left a row outside the closed form. */
int dzb_solver_block_status(dzb_solver* s, dzb_comm* c);

/* ---- CSV snapshots: Writer (writer.rs:47-146), host code ------------------------------------------------------- */
/* `f64::to_string()` of Rust: shortest round-trip digits, never an exponent, "1" for 1.0, "-0", "NaN", "inf". */
int dzb_format_f64(double v, char* buf, size_t cap);
/* Writer::new: DZB_IO unless `dir` exists; erase_prev_dir != 0 removes the plain files directly inside it. */
int dzb_csv_prepare_dir(const char* dir, int erase_prev_dir);
/* Writer::write: <dir>/<prefix><id>.csv, header = names joined by ',', then n_vars values per line. */
int dzb_csv_write(const char* dir, const char* prefix, double id, const char* const* variable_names, size_t n_vars,
                  const double* vals, size_t n_vals);

/* ---- ensembles: B independent time-dependent bars, one solve() each per step ------------------------------- */
/* meshes: host, bar-major [n_bars][n]; initial_conditions: host, [n_bars][n-2] (n_ic = n-2 per bar). */
int dzb_ensemble_td_create(const double* meshes, size_t n_bars, size_t n, double mu, double b, double bc_left,
                           double bc_right, const double* initial_conditions, size_t n_ic, size_t gauss_step,
                           const dzb_options* opt, dzb_ensemble** out);
/* `steps` successive solve(time_step) on every bar, states stay on the device */
int dzb_ensemble_step(dzb_ensemble* e, double time_step, size_t steps);
/* one solve(time_step) on every bar; out_host[n_bars][n] holds all states on return (use dzb_host_alloc memory for PCIe speed) */
int dzb_ensemble_solve(dzb_ensemble* e, double time_step, double* out_host, size_t len);
/* the same step with the states cast to f32 on the device before they cross PCIe (half the bytes): what a renderer
 * uploads anyway — Drawable::get_vertices casts every f64 to f32 for the GL buffer (mesh/mod.rs:108-113).  The device state
 * stays f64; out_host[n_bars][n] receives (float)u. */
int dzb_ensemble_solve_f32(dzb_ensemble* e, double time_step, float* out_host, size_t len);
int dzb_ensemble_state(const dzb_ensemble* e, double* out_host, size_t len);
int dzb_ensemble_read_bars(const dzb_ensemble* e, const size_t* bar_ids, size_t count, double* out_host);
int dzb_ensemble_state_device(const dzb_ensemble* e, const double** device_ptr); /* bar-major [n_bars][n] */
int dzb_ensemble_bands(const dzb_ensemble* e, size_t bar, int which, double* lo, double* di, double* up, size_t n);
int dzb_ensemble_dims(const dzb_ensemble* e, size_t* n_bars, size_t* n);
/* name of the kernel one step of this ensemble dispatches to (decided by the options, the sizes and, for the on-chip
 * operator rebuild, by the first step's check of every row): what bench.py prints as roofline.kernel */
int dzb_ensemble_step_kernel(const dzb_ensemble* e, char* buf, size_t cap);
void dzb_ensemble_destroy(dzb_ensemble* e);

/* ---- EulerSolver (reference solvers/euler/mod.rs:26-59; integration tests tests/euler.rs) for many trajectories ------ */
/* The reference steps ONE trajectory with a host closure f(&values): values = [y^(n-1), ..., y, t], `width` = 2..5 entries
 * (FunctionArguments, mod.rs:17-22).  Here every trajectory carries an affine derivative function
 *     f(values) = coeffs[0] + coeffs[1] * values[0] + ... + coeffs[width] * values[width - 1]   (left to right)
 * and do_step's arithmetic is the reference's, operation for operation.  coeffs: host [n_traj][width + 1], values: host
 * [n_traj][width].  DZB_INVALID unless 2 <= width <= 5. */
typedef struct dzb_euler dzb_euler;
int dzb_euler_create(size_t n_traj, size_t width, const double* coeffs, const double* values, dzb_euler** out);
/* Derivative functions that are not affine: the expression itself, as a postfix program shared by all trajectories — the
 * device analogue of the reference's `Fn(&[f64; N]) -> f64` (mod.rs:24-31) for closed-form expressions of the OLD values,
 * per-trajectory parameters and constants.  CONST pushes `value`, VALUE pushes values[arg], PARAM pushes this trajectory's
 * params[arg]; binary operations pop b then a and push (a op b).  IEEE double, one rounding per operation in program order,
 * no contraction: + - * / neg abs sqrt give the bits the Rust closure gives; sin cos exp log use the device's libm.  At most
 * 96 operations and 16 stack entries; DZB_INVALID for a malformed program.  params: host [n_traj][n_params]. */
typedef enum dzb_euler_opcode {
  DZB_EOP_CONST = 0, DZB_EOP_VALUE = 1, DZB_EOP_PARAM = 2, DZB_EOP_ADD = 3, DZB_EOP_SUB = 4, DZB_EOP_MUL = 5, DZB_EOP_DIV = 6,
  DZB_EOP_NEG = 7, DZB_EOP_SQRT = 8, DZB_EOP_ABS = 9, DZB_EOP_SIN = 10, DZB_EOP_COS = 11, DZB_EOP_EXP = 12, DZB_EOP_LOG = 13
} dzb_euler_opcode;
typedef struct dzb_euler_op {
  int32_t opcode; /* dzb_euler_opcode */
  int32_t arg;
  double value;
} dzb_euler_op;
int dzb_euler_create_program(size_t n_traj, size_t width, const dzb_euler_op* program, size_t n_ops, const double* params,
                             size_t n_params, const double* values, dzb_euler** out);
/* `steps` successive do_step(values, step) on every trajectory; the states stay on the device */
int dzb_euler_step(dzb_euler* e, double step, size_t steps);
int dzb_euler_set_values(dzb_euler* e, const double* values, size_t len);     /* [n_traj][width]: restart from these states */
int dzb_euler_values(const dzb_euler* e, double* out_host, size_t len);      /* [n_traj][width] */
int dzb_euler_values_device(const dzb_euler* e, const double** device_ptr);  /* SoA [width][n_traj] */
void dzb_euler_destroy(dzb_euler* e);

#ifdef __cplusplus
}
#endif
#endif /* DZAHUI_B200_H */
