/*
 * mara3_b200.h -- C ABI of the B200-native implementation of Mara3's `binary`
 * isothermal-2D hot path.  Plain pointers and sizes only; no C++ or torch types.
 *
 * The reference has no FFI layer: its seams for this path are the C++ functions
 * declared in Mara3 src/subprog_binary.hpp:180-208.  Each entry point below names
 * the reference interface it replaces.  A reference-side binding (C++ shim that
 * converts solution_t <-> flat arrays) is shown in INTEGRATION.md.
 *
 * Array layouts (all fp64, little endian, contiguous):
 *   conserved state   [B][3][N][N]   block-major; fields sigma, px, py; each block is the
 *                                    reference's row-major (N,N) array (axis 0 = x, y fastest)
 *   vertices          [B][2][N+1][N+1], cell_centers [B][2][N][N], areas / buffer rate [B][N][N]
 *   tree index        [B][3] int64   (level, i, j); blocks are in the reference's traversal
 *                                    order (depth first, child n = bx + 2 by == Morton order)
 *   scalars           43 doubles:    time, iteration numerator, denominator,
 *                                    mass_accreted_on[2], angular_momentum_accreted_on[2],
 *                                    integrated_torque_on[2], work_done_on[2], mass_ejected,
 *                                    angular_momentum_ejected, then orbital_elements_acc,
 *                                    orbital_elements_grav, orbital_elements as 10 doubles each:
 *                                    pomega, tau, cm_position_x, cm_position_y, cm_velocity_x,
 *                                    cm_velocity_y, separation, total_mass, mass_ratio, eccentricity
 *                                    (subprog_binary.hpp:108-126, model_two_body.hpp:40-62)
 *
 * With conserve_linear_p=0 the conserved state is conserved_q = (sigma, S_r, L_z) (advance_q, scheme.cpp:906-1020)
 * in the same [block][3][N][N] layout; everything else is unchanged.
 *
 * Status codes: 0 ok; 1 negative density in the updated state (the reference throws
 * std::runtime_error, scheme.cpp:747-750); 2 unbound orbit (model_two_body.hpp:385-386);
 * 3 unsupported option; -1 other error (see m3b_last_error).  There is no CPU fallback:
 * m3b_solver_create fails if no CUDA device is present.
 */
#ifndef MARA3_B200_H
#define MARA3_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct m3b_solver m3b_solver_t;       /* run_config + solver_data_t + device context */
typedef struct m3b_solution m3b_solution_t;   /* solution_t: conserved field on the GPU + scalars */

#define M3B_NUM_SCALARS 43
#define M3B_OK 0
#define M3B_NEGATIVE_DENSITY 1
#define M3B_UNBOUND_ORBIT 2
#define M3B_UNSUPPORTED 3
#define M3B_ERROR (-1)

/* ---- library ---------------------------------------------------------------------------- */
const char* m3b_version(void);
int         m3b_device_count(void);                     /* number of visible CUDA devices (0 if none) */
const char* m3b_global_error(void);                     /* message of the last failed *_create call */

/* ---- solver: create_run_config + create_solver_data + set_scheme_globals -----------------
 * (subprog_binary.cpp:155-164, subprog_binary_solver_data.cpp:18-115, scheme.cpp:42-49).
 * argv holds "key=value" tokens with the reference's 39 keys; unknown keys, duplicates and
 * bad values fail exactly as the reference's config_t does (app_config.hpp:103-136, 223-245).
 * flags: bit 0 = route every block through the general (any-tree) kernels (testing);
 *        bit 1 = host only: solver_data queries without a device context;
 *        bit 2 = use the tiled stage kernel even where the warp-strip kernel applies (testing). */
m3b_solver_t* m3b_solver_create(int argc, const char* const* argv, int device, int flags);
/* Multi-GPU: one process per GPU.  Rank 0 draws an NCCL unique id (128 bytes), hands it to the other
 * ranks by any means (bench.py uses torch.distributed), and every rank creates its solver.  The leaf
 * blocks are cut into `nranks` Morton-contiguous ranges; each rank owns one range and keeps ghost
 * copies of the remote blocks its range touches.  Per-block queries below then refer to the OWNED
 * blocks of the calling rank.  Any 2:1-balanced tree (uniform or nested): ranks keep whole ghost copies of the blocks
 * their jump blocks read. */
int         m3b_nccl_unique_id(unsigned char* out128);
m3b_solver_t* m3b_solver_create_distributed(int argc, const char* const* argv, int device, int flags,
                                            int rank, int nranks, const unsigned char* nccl_unique_id);
void        m3b_solver_destroy(m3b_solver_t* s);
const char* m3b_last_error(const m3b_solver_t* s);

/* solver_data_t queries (subprog_binary.hpp:74-104) */
int         m3b_num_blocks(const m3b_solver_t* s);              /* blocks owned by this rank (all of them on one rank) */
int         m3b_num_global_blocks(const m3b_solver_t* s);       /* leaves of the whole tree */
int         m3b_first_block(const m3b_solver_t* s);             /* global (Morton) id of the first owned block */
int         m3b_num_local_blocks(const m3b_solver_t* s);        /* owned + ghost blocks stored on this rank */
int64_t     m3b_num_owned_cells(const m3b_solver_t* s);
/* guard-zone exchange plan (host side; available without a device): local -> global block ids,
 * per peer the ordered (local block, di, dj) strips sent / received, and the 3x3 same-level
 * neighbour table [owned][9] in local ids that the stage kernel reads its halo through */
void        m3b_local_to_global(const m3b_solver_t* s, int* out);
int         m3b_halo_plan_size(const m3b_solver_t* s, int peer, int send);
void        m3b_halo_plan(const m3b_solver_t* s, int peer, int send, int* out);
void        m3b_neighbor_table(const m3b_solver_t* s, int* out);
/* [local block][4 sides: -x, +x, -y, +y][kind (0 same, 1 coarser, 2 finer), 4 leaf ids in local numbering or -1]:
 * the face-neighbour table the any-tree kernels read guard cells through (mesh_tree_operators.hpp:223-252) */
void        m3b_face_neighbor_table(const m3b_solver_t* s, int* out);
uint64_t    m3b_halo_bytes_per_exchange(const m3b_solver_t* s);
/* guard-zone transport in use: 0 none (one rank), 1 NCCL send / recv, 2 peer memory over NVLink (CUDA IPC mailboxes) */
int         m3b_exchange_transport(const m3b_solver_t* s);
/* ---- HDF5 products and the subprogram itself (no libhdf5 needed: mara3_b200/csrc/h5lite.cpp) ----
 * (with several ranks the writers are collective calls: rank 0 writes the file's structure and its own blocks, every other rank
 *  stores its blocks at their addresses in that file; if any rank fails, every rank returns M3B_ERROR)
 * m3b_write_checkpoint   replaces mara::write<state_t> into chkpt.NNNN.h5 (subprog_binary_io.cpp:131-158) for a solution
 *                        with the initial schedule and an empty time series; the run loop below stores its full state
 * m3b_write_diagnostics  replaces mara::write<diagnostic_fields_t> (subprog_binary_io.cpp:160-172, subprog_binary_diagnostics.cpp:48-82)
 * m3b_read_checkpoint    replaces mara::read<solution_t> (subprog_binary_io.cpp:174-190): the restart file must hold the solver's mesh
 * m3b_time_series_sample replaces record_time_series' sample (subprog_binary.cpp:358-379): the 47 doubles of time_series_sample_t
 * m3b_binary_main        replaces subprog_binary::main (subprog_binary.cpp:414-436): `binary key=value ...`, same stdout lines and files */
/* dataset name of a leaf in the products: mara::format_tree_index (app_serialize_tree.hpp:72-87), "level:ii-jj" */
void        m3b_format_tree_index(int level, int i, int j, char* out, int out_len);
int         m3b_write_checkpoint(m3b_solver_t* s, const m3b_solution_t* u, const char* filename);
int         m3b_write_diagnostics(m3b_solver_t* s, const m3b_solution_t* u, const char* filename);
int         m3b_read_checkpoint(m3b_solver_t* s, m3b_solution_t* u, const char* filename);
int         m3b_time_series_sample(m3b_solver_t* s, const m3b_solution_t* u, double* out47);
int         m3b_binary_main(int argc, const char* const* argv, int device);
/* the same on `nranks` GPUs, one process each: every rank runs the loop, rank 0 prints and writes the gathered products */
int         m3b_binary_main_distributed(int argc, const char* const* argv, int device, int rank, int nranks, const unsigned char* nccl_unique_id128);
/* test hook of the built-in HDF5 writer / reader (h5lite): write one file with every structure it emits and / or
 * list the root group of an existing file into `report`; 0 or -1 (message in `report`) */
int         m3b_h5_selftest(const char* write_path, const char* read_path, char* report, int report_len);
/* the same content written by one process into `whole_path` and by `parts` processes' roles (one after another) into
 * `shared_path`: the files must be byte-identical (how the product writers share a file between ranks); 0 or -1 */
int         m3b_h5_selftest_shared(const char* whole_path, const char* shared_path, int parts);
int         m3b_block_size(const m3b_solver_t* s);
int64_t     m3b_num_cells(const m3b_solver_t* s);
int         m3b_num_regular_blocks(const m3b_solver_t* s);   /* blocks served by the fused kernel */
void        m3b_tree_index(const m3b_solver_t* s, int64_t* out);
void        m3b_vertices(const m3b_solver_t* s, double* out);
void        m3b_cell_centers(const m3b_solver_t* s, double* out);
void        m3b_cell_areas(const m3b_solver_t* s, double* out);
void        m3b_buffer_rate_field(const m3b_solver_t* s, double* out);
void        m3b_initial_conserved_u(const m3b_solver_t* s, double* out);
double      m3b_recommended_time_step(const m3b_solver_t* s);
double      m3b_gst_suppr_radius(const m3b_solver_t* s);
double      m3b_density_floor(const m3b_solver_t* s);
/* value of a run_config key as text (pretty_print formatting); returns 0, or 1 if the key is unknown */
int         m3b_config_get(const m3b_solver_t* s, const char* key, char* out, int out_len);

/* ---- solution: create_solution (subprog_binary.cpp:196-227) and value-type plumbing ------- */
m3b_solution_t* m3b_solution_create(m3b_solver_t* s);                 /* initial disk */
m3b_solution_t* m3b_solution_clone(m3b_solver_t* s, const m3b_solution_t* u);
void        m3b_solution_destroy(m3b_solution_t* u);
int         m3b_solution_set_conserved(m3b_solver_t* s, m3b_solution_t* u, const double* host);   /* H2D */
int         m3b_solution_get_conserved(m3b_solver_t* s, const m3b_solution_t* u, double* host);   /* D2H */
void        m3b_solution_set_scalars(m3b_solution_t* u, const double* in43);
void        m3b_solution_get_scalars(const m3b_solution_t* u, double* out43);

/* ---- operators ---------------------------------------------------------------------------- */
/* binary::maximum_timestep (scheme.cpp:1107-1126) */
int         m3b_maximum_timestep(m3b_solver_t* s, const m3b_solution_t* u, double* dt_out);
/* binary::advance (scheme.cpp:1022-1027): out = one RK stage of `in`; `in` is unchanged */
int         m3b_advance(m3b_solver_t* s, const m3b_solution_t* in, double dt, int safe_mode, m3b_solution_t* out);
/* solution_t::operator+ / operator* (scheme.cpp:1033-1069): out = a * b0 + b * (1 - b0) */
int         m3b_solution_combine(m3b_solver_t* s, const m3b_solution_t* a, const m3b_solution_t* b, double b0, m3b_solution_t* out);
/* binary::next_solution (subprog_binary.cpp:258-293), in place: dt rule, RK1 / RK2, safe-mode retry */
int         m3b_next_solution(m3b_solver_t* s, m3b_solution_t* u, double* dt_used, int* fell_back);
/* `count` calls of m3b_next_solution; stops at the first non-zero status */
int         m3b_run_steps(m3b_solver_t* s, m3b_solution_t* u, int count, int* num_fallbacks);

/* Host-buffer drop-ins: the same two operators with the state in caller memory
 * (conserved [B][3][N][N] + 43 scalars in, the same out); copies are part of the call. */
int         m3b_advance_host(m3b_solver_t* s, const double* u_in, const double* scalars_in, double dt, int safe_mode,
                             double* u_out, double* scalars_out);
int         m3b_next_solution_host(m3b_solver_t* s, const double* u_in, const double* scalars_in,
                                   double* u_out, double* scalars_out, double* dt_used, int* fell_back);

/* ---- two-body model (host only; model_two_body.hpp:220-281, 295-402) ------------------------
 * elements: 10 doubles as in the scalars; bodies: mass, x, y, vx, vy of body 1 then body 2 */
int         m3b_two_body_state(const double* elements10, double t, double* bodies10_out);
int         m3b_orbital_elements(const double* bodies10, double t, double* elements10_out);   /* 0, or 2 if unbound */

/* ---- diagnostics --------------------------------------------------------------------------- */
/* lines the reference prints before throwing ("negative density ... (at position [...])") */
int         m3b_num_messages(const m3b_solver_t* s);
const char* m3b_message(const m3b_solver_t* s, int n);
void        m3b_set_quiet(m3b_solver_t* s, int quiet);                /* 1: do not print those lines to stdout */
/* m3b_next_solution queues the FOLLOWING step on the GPU (dt and body positions computed on the device)
 * before it waits for the current one, so that a stepping loop never idles the GPU on the host.
 * 0 switches this off (every call then starts and finishes exactly one step). Default 1. */
void        m3b_set_pipelining(m3b_solver_t* s, int on);
uint64_t    m3b_kernel_launches(const m3b_solver_t* s);               /* kernels launched so far */
/* CUDA-event timing of the fused stage kernel on its own stream (for the roofline) */
void        m3b_stage_timing(m3b_solver_t* s, int enable);
int         m3b_stage_timing_read(m3b_solver_t* s, double* total_ms, uint64_t* launches);
/* Several ranks, with stage timing on: out8 = { steps instrumented, us the compute stream waited for ghost blocks (interior
 * blocks done -> blocks with ghost neighbours may start), us the step's last finish kernel waited for the other ranks'
 * results, us from "stage input ready" to "ghost blocks unpacked" on the exchange stream, exchanges, bytes pushed,
 * us CTAs of the stage kernel spent in its fused unpack (flag wait + their share of the scatter), number of such visits },
 * totals since the last call.  M3B_ERROR on one rank.  (The reference has no counterpart: subprog_binary_scheme.cpp:132-142
 * reads its neighbours through shared memory.) */
int         m3b_exchange_timing(m3b_solver_t* s, double* out8);
/* launch on a caller-owned cudaStream_t from now on (NULL: back to the solver's own stream) */
int         m3b_set_stream(m3b_solver_t* s, void* cuda_stream);
void        m3b_synchronize(m3b_solver_t* s);

#ifdef __cplusplus
}
#endif
#endif
