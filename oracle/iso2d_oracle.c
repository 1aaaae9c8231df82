/*
 * oracle/iso2d_oracle.c -- TEST INFRASTRUCTURE ONLY (parity checker).
 *
 * Plain-C, single-threaded re-statement of the reference's `binary` iso2d hot
 * path.  See iso2d_oracle.h for the pinning statement.  Every function cites the
 * reference file:line it follows (paths are relative to Mara3's src/).  The
 * floating-point operation ORDER of the reference is kept on purpose (compile
 * with -ffp-contract=off), so on the same libm this code is expected to be
 * bit-identical to the reference's parity build (g++ -O2, no -march).
 *
 * Block arrays are structure-of-arrays [field][i][j] with j (y) fastest, i.e. the
 * reference's row-major (N,N) block (core_ndarray.hpp:777-792) split per tuple
 * component.  Leaves are numbered in the reference's traversal order: depth
 * first, children n = bx + 2*by (core_tree.hpp:156-159, 334-337).
 */
#define _GNU_SOURCE
#include "iso2d_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif




/* ===========================================================================
 * Config (subprog_binary.cpp:57-99; app_config.hpp:103-136)
 * ======================================================================== */
void m3o_config_default(m3o_config_t* c)
{
    memset(c, 0, sizeof(*c));
    strcpy(c->outdir, "data");
    c->cpi = 10.0; c->dfi = 1.0; c->tsi = 2e-3; c->tfinal = 1.0; c->cfl_number = 0.4;
    c->fixed_dt = 0; c->depth = 4; c->begin_live_binary = 1e6; c->conserve_linear_p = 1;
    c->block_size = 24; c->focus_factor = 2.0; c->focus_index = 2.0; c->threaded = 1; c->rk_order = 2;
    strcpy(c->reconstruct_method, "plm");
    c->plm_theta = 1.8; c->source_term_softening = 1.0; c->softening_radius = 0.05;
    c->sink_radius = 0.05; c->sink_rate = 1.0; c->buffer_damping_rate = 10.0; c->domain_radius = 12.0;
    c->disk_radius = 2.0; c->disk_mass = 1e-3; c->ambient_density = 1e-4; c->density_floor = 0.0;
    c->separation = 1.0; c->mass_ratio = 1.0; c->eccentricity = 0.0; c->counter_rotate = 0;
    c->mach_number = 10.0; c->axisymmetric_cs2 = 0; c->no_accretion_force = 0;
    c->alpha_cutoff_radius = 0.0; c->alpha = 0.1; c->nu = 0.0; c->mdot = 0.0;
}

int m3o_config_set(m3o_config_t* c, const char* key, const char* value)
{
#define KEY_D(name) if (! strcmp(key, #name)) { char* e; c->name = strtod(value, &e); return e == value ? 2 : 0; }
#define KEY_I(name) if (! strcmp(key, #name)) { char* e; c->name = (int) strtol(value, &e, 10); return e == value ? 2 : 0; }
#define KEY_S(name) if (! strcmp(key, #name)) { strncpy(c->name, value, sizeof(c->name) - 1); return 0; }
    KEY_S(restart) KEY_S(outdir) KEY_D(cpi) KEY_D(dfi) KEY_D(tsi) KEY_D(tfinal) KEY_D(cfl_number)
    KEY_I(fixed_dt) KEY_I(depth) KEY_D(begin_live_binary) KEY_I(conserve_linear_p) KEY_I(block_size)
    KEY_D(focus_factor) KEY_D(focus_index) KEY_I(threaded) KEY_I(rk_order) KEY_S(reconstruct_method)
    KEY_D(plm_theta) KEY_D(source_term_softening) KEY_D(softening_radius) KEY_D(sink_radius) KEY_D(sink_rate)
    KEY_D(buffer_damping_rate) KEY_D(domain_radius) KEY_D(disk_radius) KEY_D(disk_mass) KEY_D(ambient_density)
    KEY_D(density_floor) KEY_D(separation) KEY_D(mass_ratio) KEY_D(eccentricity) KEY_I(counter_rotate)
    KEY_D(mach_number) KEY_I(axisymmetric_cs2) KEY_I(no_accretion_force) KEY_D(alpha_cutoff_radius)
    KEY_D(alpha) KEY_D(nu) KEY_D(mdot)
#undef KEY_D
#undef KEY_I
#undef KEY_S
    return 1;
}




/* ===========================================================================
 * Quadtree (core_tree.hpp:86-219, 225-912; mesh_tree_operators.hpp:90-198)
 * ======================================================================== */
typedef struct node
{
    struct node* child[4];  /* all NULL for a leaf */
    double* verts;          /* leaf: [2][N+1][N+1] vertex block */
    int leaf_id;            /* leaf: position in the traversal order */
} node_t;

struct m3o_mesh
{
    m3o_config_t config;
    int N, B;
    node_t* root;
    node_t** leaf;       /* [B] */
    int64_t* index;      /* [B][3] */
    double* vertices;    /* [B][2][N+1][N+1] */
    double* centers;     /* [B][2][N][N] */
    double* areas;       /* [B][N][N] */
    double* buffer_rate; /* [B][N][N] */
    double* U0;          /* [B][3][N][N] */
    double recommended_time_step, gst_suppr_radius, density_floor;
};

struct m3o_solution
{
    int B, N;
    double time;
    int iter_num, iter_den;
    double* U;           /* [B][3][N][N] */
    double mass_accreted_on[2], angular_momentum_accreted_on[2], integrated_torque_on[2], work_done_on[2];
    double mass_ejected, angular_momentum_ejected;
    m3o_elements_t elements_acc, elements_grav, elements;
};

static int is_leaf(const node_t* n) { return n->child[0] == NULL; }

static node_t* node_new(void)
{
    return (node_t*) calloc(1, sizeof(node_t));
}

static void node_free(node_t* n)
{
    if (! n) return;
    for (int c = 0; c < 4; ++c) node_free(n->child[c]);
    free(n->verts);
    free(n);
}

/* tree.depth(): core_tree.hpp:261-264 */
static int node_depth(const node_t* n)
{
    if (is_leaf(n)) return 0;
    int d = 0;
    for (int c = 0; c < 4; ++c) { int e = 1 + node_depth(n->child[c]); if (e > d) d = e; }
    return d;
}

/* The node (leaf or not) at exactly (level, i, j), or NULL: contains_node / node_at
 * (core_tree.hpp:426-437).  Bit (level-1) of the coordinate is the orthant at the
 * root (core_tree.hpp:190-193, 111-114). */
static node_t* find_node(node_t* root, int level, long i, long j)
{
    node_t* n = root;
    for (int l = level - 1; l >= 0; --l)
    {
        if (is_leaf(n)) return NULL;
        n = n->child[((i >> l) & 1) + 2 * ((j >> l) & 1)];
    }
    return n;
}

static node_t* find_leaf(node_t* root, int level, long i, long j)
{
    if (level < 0) return NULL;
    node_t* n = find_node(root, level, i, j);
    return (n && is_leaf(n)) ? n : NULL;
}

/* refine_verts<2> (mesh_prolong_restrict.hpp:148-159, 311-322): prolong on axis 0
 * then on axis 1 by midpoint averaging `(a + b) * 0.5`, then bisect. */
static void bifurcate(node_t* n, int N)
{
    int V = N + 1, W = 2 * N + 1;
    double* P0 = (double*) malloc(sizeof(double) * 2 * W * V);
    double* P1 = (double*) malloc(sizeof(double) * 2 * W * W);

    for (int c = 0; c < 2; ++c)
    {
        const double* C = n->verts + c * V * V;
        for (int i = 0; i < W; ++i)
            for (int j = 0; j < V; ++j)
            {
                int lo = i / 2, hi = i / 2 + (i % 2 == 0 ? 0 : 1);
                P0[(c * W + i) * V + j] = (C[lo * V + j] + C[hi * V + j]) * 0.5;
            }
        for (int i = 0; i < W; ++i)
            for (int j = 0; j < W; ++j)
            {
                int lo = j / 2, hi = j / 2 + (j % 2 == 0 ? 0 : 1);
                P1[(c * W + i) * W + j] = (P0[(c * W + i) * V + lo] + P0[(c * W + i) * V + hi]) * 0.5;
            }
    }
    for (int k = 0; k < 4; ++k)
    {
        int bx = k & 1, by = k >> 1;
        node_t* ch = node_new();
        ch->verts = (double*) malloc(sizeof(double) * 2 * V * V);
        for (int c = 0; c < 2; ++c)
            for (int i = 0; i < V; ++i)
                for (int j = 0; j < V; ++j)
                    ch->verts[(c * V + i) * V + j] = P1[(c * W + bx * N + i) * W + by * N + j];
        n->child[k] = ch;
    }
    free(n->verts);
    n->verts = NULL;
    free(P0);
    free(P1);
}

/* One pass of create_vertex_quadtree's loop (mesh_tree_operators.hpp:175-189): every
 * CURRENT leaf is tested once; `pass` (not the node's level) feeds the predicate
 * (subprog_binary.cpp:174-177). */
static void refine_pass(node_t* n, int N, int pass, double focus_factor, double focus_index)
{
    if (! is_leaf(n))
    {
        for (int c = 0; c < 4; ++c) refine_pass(n->child[c], N, pass, focus_factor, focus_index);
        return;
    }
    int V = N + 1;
    double cx = (n->verts[0] + n->verts[N * V + N]) * 0.5;
    double cy = (n->verts[V * V] + n->verts[V * V + N * V + N]) * 0.5;
    double centroid_radius = sqrt(cx * cx + cy * cy);

    if (centroid_radius < focus_factor / pow((double) pass, focus_index))
    {
        bifurcate(n, N);
    }
}

/* over_refined_neighbors (mesh_tree_operators.hpp:90-100) for one leaf */
static int over_refined(node_t* root, int level, long i, long j)
{
    long n = 1L << level;
    long ni[4] = {(i + n + 1) % n, (i + n - 1) % n, i, i};
    long nj[4] = {j, j, (j + n + 1) % n, (j + n - 1) % n};
    for (int k = 0; k < 4; ++k)
    {
        node_t* nb = find_node(root, level, ni[k], nj[k]);
        if (nb && node_depth(nb) > 1) return 1;
    }
    return 0;
}

static void collect_over_refined(node_t* root, node_t* n, int level, long i, long j, node_t*** list, int* count, int* cap)
{
    if (is_leaf(n))
    {
        if (over_refined(root, level, i, j))
        {
            if (*count == *cap) { *cap = *cap ? 2 * *cap : 64; *list = (node_t**) realloc(*list, sizeof(node_t*) * *cap); }
            (*list)[(*count)++] = n;
        }
        return;
    }
    for (int c = 0; c < 4; ++c)
        collect_over_refined(root, n->child[c], level + 1, 2 * i + (c & 1), 2 * j + (c >> 1), list, count, cap);
}

/* ensure_valid_quadtree (mesh_tree_operators.hpp:115-139): flag on the current tree,
 * split all flagged leaves at once, repeat until none is flagged. */
static void ensure_valid(node_t* root, int N)
{
    for (;;)
    {
        node_t** list = NULL; int count = 0, cap = 0;
        collect_over_refined(root, root, 0, 0, 0, &list, &count, &cap);
        for (int k = 0; k < count; ++k) bifurcate(list[k], N);
        free(list);
        if (count == 0) break;
    }
}

static void count_leaves(node_t* n, int* count)
{
    if (is_leaf(n)) { n->leaf_id = (*count)++; return; }
    for (int c = 0; c < 4; ++c) count_leaves(n->child[c], count);
}

static void fill_leaves(node_t* n, int level, long i, long j, m3o_mesh_t* m)
{
    if (is_leaf(n))
    {
        m->leaf[n->leaf_id] = n;
        m->index[3 * n->leaf_id + 0] = level;
        m->index[3 * n->leaf_id + 1] = i;
        m->index[3 * n->leaf_id + 2] = j;
        return;
    }
    for (int c = 0; c < 4; ++c) fill_leaves(n->child[c], level + 1, 2 * i + (c & 1), 2 * j + (c >> 1), m);
}




/* ===========================================================================
 * Initial disk model (subprog_binary.cpp:105-153)
 * ======================================================================== */
static double disk_sigma(const m3o_config_t* c, double r)
{
    double rc = c->disk_radius;
    double s0 = c->disk_mass / (17.0618 * rc * rc);
    double s1 = c->ambient_density * s0;
    double x = r / rc;
    return s0 * exp(-0.5 * (x - 1) * (x - 1)) + s1;
}

static void disk_profile(const m3o_config_t* c, double x, double y, double* p)
{
    double rc = c->disk_radius;
    double s0 = c->disk_mass / (17.0618 * rc * rc);
    double s1 = c->ambient_density * s0;
    double r2 = x * x + y * y;
    double r  = sqrt(r2);
    double rs = c->softening_radius;
    double GM = 1.0;
    double Ma = c->mach_number;
    double xx = r / rc;
    double dp_dr = (GM / Ma / Ma / (r + rs)) * (xx * (1 - xx) * (1 - s1 / disk_sigma(c, r)) - 1.0);
    double vp = sqrt(GM / (r + rs) + dp_dr) * (c->counter_rotate ? -1 : 1);
    double vr = -c->mdot / (disk_sigma(c, r) * 2 * M_PI * r) * (r > 2.0);
    p[0] = disk_sigma(c, r);
    p[1] = vr * (x / r) + vp * (-y / r);
    p[2] = vr * (y / r) + vp * ( x / r);
}




/* ===========================================================================
 * Mesh + solver_data (subprog_binary.cpp:166-184; subprog_binary_solver_data.cpp:18-115)
 * ======================================================================== */
m3o_mesh_t* m3o_mesh_create(const m3o_config_t* c)
{
    m3o_mesh_t* m = (m3o_mesh_t*) calloc(1, sizeof(m3o_mesh_t));
    int N = c->block_size, V = N + 1;
    m->config = *c;
    m->N = N;

    /* root block: cartesian product of linspace(-1, 1, N+1) (core_ndarray.hpp:2544-2551) */
    m->root = node_new();
    m->root->verts = (double*) malloc(sizeof(double) * 2 * V * V);
    for (int i = 0; i < V; ++i)
        for (int j = 0; j < V; ++j)
        {
            m->root->verts[i * V + j]         = -1.0 + (1.0 - -1.0) * i / (V - 1);
            m->root->verts[V * V + i * V + j] = -1.0 + (1.0 - -1.0) * j / (V - 1);
        }
    for (int pass = 0; pass < c->depth; ++pass)
        refine_pass(m->root, N, pass, c->focus_factor, c->focus_index);
    ensure_valid(m->root, N);

    count_leaves(m->root, &m->B);
    int B = m->B;
    m->leaf  = (node_t**) malloc(sizeof(node_t*) * B);
    m->index = (int64_t*) malloc(sizeof(int64_t) * 3 * B);
    fill_leaves(m->root, 0, 0, 0, m);

    m->vertices    = (double*) malloc(sizeof(double) * (size_t) B * 2 * V * V);
    m->centers     = (double*) malloc(sizeof(double) * (size_t) B * 2 * N * N);
    m->areas       = (double*) malloc(sizeof(double) * (size_t) B * N * N);
    m->buffer_rate = (double*) malloc(sizeof(double) * (size_t) B * N * N);
    m->U0          = (double*) malloc(sizeof(double) * (size_t) B * 3 * N * N);

    double min_dx = INFINITY, min_dy = INFINITY, max_v = 0.0;

    for (int b = 0; b < B; ++b)
    {
        double* xv = m->vertices + (size_t) b * 2 * V * V;
        double* yv = xv + V * V;
        double* xc = m->centers + (size_t) b * 2 * N * N;
        double* yc = xc + N * N;
        double* dA = m->areas + (size_t) b * N * N;
        double* br = m->buffer_rate + (size_t) b * N * N;
        double* U0 = m->U0 + (size_t) b * 3 * N * N;

        /* (block * domain_radius) : subprog_binary.cpp:180-183 */
        for (int k = 0; k < 2 * V * V; ++k) xv[k] = m->leaf[b]->verts[k] * c->domain_radius;

        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
            {
                /* midpoint_on_axis(0) then (1): solver_data.cpp:22-27, core_ndarray_ops.hpp:121-129 */
                double mx0 = (xv[i * V + j]     + xv[(i + 1) * V + j])     * 0.5;
                double mx1 = (xv[i * V + j + 1] + xv[(i + 1) * V + j + 1]) * 0.5;
                double my0 = (yv[i * V + j]     + yv[(i + 1) * V + j])     * 0.5;
                double my1 = (yv[i * V + j + 1] + yv[(i + 1) * V + j + 1]) * 0.5;
                double x = (mx0 + mx1) * 0.5;
                double y = (my0 + my1) * 0.5;
                xc[i * N + j] = x;
                yc[i * N + j] = y;

                /* cell areas: solver_data.cpp:29-34 */
                double dx0 = xv[(i + 1) * V + j]     - xv[i * V + j];
                double dx1 = xv[(i + 1) * V + j + 1] - xv[i * V + j + 1];
                double dy0 = yv[i * V + j + 1]       - yv[i * V + j];
                double dy1 = yv[(i + 1) * V + j + 1] - yv[(i + 1) * V + j];
                dA[i * N + j] = ((dx0 + dx1) * 0.5) * ((dy0 + dy1) * 0.5);

                /* buffer_rate_field: solver_data.cpp:64-78 */
                double rc = pow(x * x + y * y, 0.5);
                double yy = 3.0 * (rc - c->domain_radius);
                br[i * N + j] = c->buffer_damping_rate * (1.0 + tanh(yy));

                /* initial conserved: subprog_binary.cpp:198-205, physics_iso2d.hpp:249-258 */
                double p[3];
                disk_profile(c, x, y, p);
                if (c->conserve_linear_p)
                {
                    U0[0 * N * N + i * N + j] = p[0];
                    U0[1 * N * N + i * N + j] = p[0] * p[1];
                    U0[2 * N * N + i * N + j] = p[0] * p[2];
                }
                else    /* to_conserved_angmom_per_area (physics_iso2d.hpp:263-272): (sigma, Sr, Lz); held in the same array */
                {
                    U0[0 * N * N + i * N + j] = p[0];
                    U0[1 * N * N + i * N + j] = p[0] * (x * p[1] + y * p[2]);
                    U0[2 * N * N + i * N + j] = p[0] * (x * p[2] - y * p[1]);
                }

                double vmag = sqrt(p[1] * p[1] + p[2] * p[2]);
                if (vmag > max_v) max_v = vmag;
            }
        /* min_dx / min_dy over ALL vertex differences: solver_data.cpp:41-55 */
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < V; ++j)
            {
                double d = xv[(i + 1) * V + j] - xv[i * V + j];
                if (d < min_dx) min_dx = d;
            }
        for (int i = 0; i < V; ++i)
            for (int j = 0; j < N; ++j)
            {
                double d = yv[i * V + j + 1] - yv[i * V + j];
                if (d < min_dy) min_dy = d;
            }
    }
    double max_velocity = max_v > 1.0 ? max_v : 1.0;              /* solver_data.cpp:57-62 */
    double min_d = min_dx < min_dy ? min_dx : min_dy;
    m->gst_suppr_radius      = c->source_term_softening * min_d;   /* :91 */
    m->density_floor         = c->density_floor * c->disk_mass;    /* :100 */
    m->recommended_time_step = min_d / max_velocity * c->cfl_number; /* :102 */
    return m;
}

void m3o_mesh_destroy(m3o_mesh_t* m)
{
    if (! m) return;
    node_free(m->root);
    free(m->leaf); free(m->index); free(m->vertices); free(m->centers);
    free(m->areas); free(m->buffer_rate); free(m->U0);
    free(m);
}

int    m3o_mesh_num_blocks(const m3o_mesh_t* m) { return m->B; }
int    m3o_mesh_block_size(const m3o_mesh_t* m) { return m->N; }
void   m3o_mesh_tree_index(const m3o_mesh_t* m, int64_t* out) { memcpy(out, m->index, sizeof(int64_t) * 3 * m->B); }
void   m3o_mesh_vertices(const m3o_mesh_t* m, double* out) { memcpy(out, m->vertices, sizeof(double) * (size_t) m->B * 2 * (m->N + 1) * (m->N + 1)); }
void   m3o_mesh_cell_centers(const m3o_mesh_t* m, double* out) { memcpy(out, m->centers, sizeof(double) * (size_t) m->B * 2 * m->N * m->N); }
void   m3o_mesh_cell_areas(const m3o_mesh_t* m, double* out) { memcpy(out, m->areas, sizeof(double) * (size_t) m->B * m->N * m->N); }
void   m3o_mesh_buffer_rate(const m3o_mesh_t* m, double* out) { memcpy(out, m->buffer_rate, sizeof(double) * (size_t) m->B * m->N * m->N); }
void   m3o_mesh_initial_conserved_u(const m3o_mesh_t* m, double* out) { memcpy(out, m->U0, sizeof(double) * (size_t) m->B * 3 * m->N * m->N); }
double m3o_mesh_recommended_time_step(const m3o_mesh_t* m) { return m->recommended_time_step; }
double m3o_mesh_gst_suppr_radius(const m3o_mesh_t* m) { return m->gst_suppr_radius; }
double m3o_mesh_density_floor(const m3o_mesh_t* m) { return m->density_floor; }




/* ===========================================================================
 * Solution (subprog_binary.hpp:108-126; subprog_binary.cpp:196-227)
 * ======================================================================== */
static m3o_elements_t elements_zeros(void) { m3o_elements_t e; memset(&e, 0, sizeof(e)); return e; }

m3o_solution_t* m3o_solution_create(const m3o_mesh_t* m)
{
    m3o_solution_t* s = (m3o_solution_t*) calloc(1, sizeof(m3o_solution_t));
    size_t n = (size_t) m->B * 3 * m->N * m->N;
    s->B = m->B; s->N = m->N;
    s->time = 0.0; s->iter_num = 0; s->iter_den = 1;
    s->U = (double*) malloc(sizeof(double) * n);
    memcpy(s->U, m->U0, sizeof(double) * n);
    s->elements_acc  = elements_zeros();           /* make_full_orbital_elements_with_zeros */
    s->elements_grav = elements_zeros();
    s->elements      = elements_zeros();           /* make_full_orbital_elements(create_binary_params) */
    s->elements.v[6] = m->config.separation;
    s->elements.v[7] = 1.0;
    s->elements.v[8] = m->config.mass_ratio;
    s->elements.v[9] = m->config.eccentricity;
    return s;
}

m3o_solution_t* m3o_solution_clone(const m3o_solution_t* s)
{
    m3o_solution_t* r = (m3o_solution_t*) malloc(sizeof(m3o_solution_t));
    size_t n = (size_t) s->B * 3 * s->N * s->N;
    *r = *s;
    r->U = (double*) malloc(sizeof(double) * n);
    memcpy(r->U, s->U, sizeof(double) * n);
    return r;
}

static void solution_assign(m3o_solution_t* dst, const m3o_solution_t* src)
{
    double* U = dst->U;
    size_t n = (size_t) src->B * 3 * src->N * src->N;
    *dst = *src;
    dst->U = U;
    memcpy(dst->U, src->U, sizeof(double) * n);
}

void m3o_solution_destroy(m3o_solution_t* s) { if (s) { free(s->U); free(s); } }
void m3o_solution_get_conserved(const m3o_solution_t* s, double* out) { memcpy(out, s->U, sizeof(double) * (size_t) s->B * 3 * s->N * s->N); }
void m3o_solution_set_conserved(m3o_solution_t* s, const double* in) { memcpy(s->U, in, sizeof(double) * (size_t) s->B * 3 * s->N * s->N); }

void m3o_solution_get_scalars(const m3o_solution_t* s, double* o)
{
    o[0] = s->time; o[1] = s->iter_num; o[2] = s->iter_den;
    for (int k = 0; k < 2; ++k)
    {
        o[3 + k] = s->mass_accreted_on[k]; o[5 + k] = s->angular_momentum_accreted_on[k];
        o[7 + k] = s->integrated_torque_on[k]; o[9 + k] = s->work_done_on[k];
    }
    o[11] = s->mass_ejected; o[12] = s->angular_momentum_ejected;
    memcpy(o + 13, s->elements_acc.v, sizeof(double) * 10);
    memcpy(o + 23, s->elements_grav.v, sizeof(double) * 10);
    memcpy(o + 33, s->elements.v, sizeof(double) * 10);
}

void m3o_solution_set_scalars(m3o_solution_t* s, const double* o)
{
    s->time = o[0]; s->iter_num = (int) o[1]; s->iter_den = (int) o[2];
    for (int k = 0; k < 2; ++k)
    {
        s->mass_accreted_on[k] = o[3 + k]; s->angular_momentum_accreted_on[k] = o[5 + k];
        s->integrated_torque_on[k] = o[7 + k]; s->work_done_on[k] = o[9 + k];
    }
    s->mass_ejected = o[11]; s->angular_momentum_ejected = o[12];
    memcpy(s->elements_acc.v, o + 13, sizeof(double) * 10);
    memcpy(s->elements_grav.v, o + 23, sizeof(double) * 10);
    memcpy(s->elements.v, o + 33, sizeof(double) * 10);
}




/* ===========================================================================
 * Two-body model (model_two_body.hpp)
 * ======================================================================== */
static double orbital_period(const double* el)  /* :424-429 */
{
    double M = el[7], a = el[6];
    return 2 * M_PI / sqrt(M / a / a / a);
}

/* compute_two_body_state(orbital_elements_t, t): model_two_body.hpp:168-218 */
static void two_body_local(const double* el, double t, m3o_two_body_t* r)
{
    double e = el[9], q = el[8], a = el[6], M = el[7];
    double omega = a == 0.0 ? 0.0 : sqrt(M / a / a / a);
    double mu = q / (1.0 + q);
    double E;

    if (e > 0.0)
    {
        double Mn = omega * t;      /* Newton-Raphson on Kepler's equation, tolerance 1e-10 (:119-135) */
        double x = Mn;
        double y = x - e * sin(x) - Mn;
        while (fabs(y) > 1e-10)
        {
            x -= y / (1 - e * cos(x));
            y = x - e * sin(x) - Mn;
        }
        E = x;
    }
    else
    {
        E = omega * t;
    }
    r->b[0][0] = M * (1 - mu);
    r->b[1][0] = M * mu;
    r->b[0][1] = -a * mu * (e - cos(E));
    r->b[0][2] = +a * mu * (0 + sin(E)) * sqrt(1 - e * e);
    r->b[1][1] = -r->b[0][1] / q;
    r->b[1][2] = -r->b[0][2] / q;
    r->b[0][3] = -a * mu * omega / (1 - e * cos(E)) * sin(E);
    r->b[0][4] = +a * mu * omega / (1 - e * cos(E)) * cos(E) * sqrt(1 - e * e);
    r->b[1][3] = -r->b[0][3] / q;
    r->b[1][4] = -r->b[0][4] / q;
}

/* compute_two_body_state(full_orbital_elements_t, t): model_two_body.hpp:220-281 */
void m3o_two_body_state(const m3o_elements_t* P, double t, m3o_two_body_t* out)
{
    const double* p = P->v;
    while (t < p[1]) t += orbital_period(p);

    m3o_two_body_t L;
    two_body_local(p, t - p[1], &L);
    double c = cos(-p[0]);
    double s = sin(-p[0]);

    for (int k = 0; k < 2; ++k)
    {
        double x = L.b[k][1], y = L.b[k][2], vx = L.b[k][3], vy = L.b[k][4];
        out->b[k][0] = L.b[k][0];
        out->b[k][1] = (+x * c + y * s) + p[2];
        out->b[k][2] = (-x * s + y * c) + p[3];
        out->b[k][3] = (+vx * c + vy * s) + p[4];
        out->b[k][4] = (-vx * s + vy * c) + p[5];
    }
}

static double clampd(double x0, double x1, double x) { return fmin(fmax(x, x0), x1); }

/* compute_orbital_elements: model_two_body.hpp:295-402 */
int m3o_orbital_elements(const m3o_two_body_t* s, double t, m3o_elements_t* out)
{
    const double* c1 = s->b[0];
    const double* c2 = s->b[1];
    double M1 = c1[0], M2 = c2[0], M = M1 + M2, q = M2 / M1;
    double x_cm  = (c1[1] * c1[0] + c2[1] * c2[0]) / M;
    double y_cm  = (c1[2] * c1[0] + c2[2] * c2[0]) / M;
    double vx_cm = (c1[3] * c1[0] + c2[3] * c2[0]) / M;
    double vy_cm = (c1[4] * c1[0] + c2[4] * c2[0]) / M;
    double x1 = c1[1] - x_cm, y1 = c1[2] - y_cm, x2 = c2[1] - x_cm, y2 = c2[2] - y_cm;
    double r1 = sqrt(x1 * x1 + y1 * y1);
    double r2 = sqrt(x2 * x2 + y2 * y2);
    double vx1 = c1[3] - vx_cm, vy1 = c1[4] - vy_cm, vx2 = c2[3] - vx_cm, vy2 = c2[4] - vy_cm;
    double vf1 = -vx1 * y1 / r1 + vy1 * x1 / r1;
    double vf2 = -vx2 * y2 / r2 + vy2 * x2 / r2;
    double v1 = sqrt(vx1 * vx1 + vy1 * vy1);
    double E1 = 0.5 * M1 * (vx1 * vx1 + vy1 * vy1);
    double E2 = 0.5 * M2 * (vx2 * vx2 + vy2 * vy2);
    double L1 = M1 * r1 * vf1;
    double L2 = M2 * r2 * vf2;
    double R = r1 + r2;
    double L = L1 + L2;
    double E = E1 + E2 - M1 * M2 / R;
    double a = -0.5 * M1 * M2 / E;
    double b = sqrt(-0.5 * L * L / E * (M1 + M2) / (M1 * M2));
    double e = sqrt(clampd(0.0, 1.0, 1.0 - b * b / a / a));
    double omega = sqrt(M / a / a / a);
    double a1 = a * q / (1.0 + q);
    double b1 = b * q / (1.0 + q);
    double cn = e == 0.0 ? x1 / r1 : (1.0 - r1 / a1) / e;
    double cf = a1 / r1 * (cn - e);
    double sn = e == 0.0 ? y1 / r1 : (vx1 * x1 + vy1 * y1) / (e * v1 * r1) * sqrt(1.0 - e * e * cn * cn);
    double sf = (b1 / r1) * sn;
    double cE = (e + cf)                / (1.0 + e * cf);
    double sE = sqrt(1.0 - e * e) * sf  / (1.0 + e * cf);
    double EE = atan2(sE, cE);
    double MM = EE - e * sE;
    double tau = t - MM / omega;
    double ax = +(cn - e) * x1 + sn * sqrt(1.0 - e * e) * y1;
    double ay = +(cn - e) * y1 - sn * sqrt(1.0 - e * e) * x1;
    double pomega = atan2(ay, ax);

    if (E >= 0.0) return 1;

    out->v[0] = pomega; out->v[1] = tau; out->v[2] = x_cm; out->v[3] = y_cm; out->v[4] = vx_cm; out->v[5] = vy_cm;
    out->v[6] = a; out->v[7] = M; out->v[8] = q; out->v[9] = e;
    return 0;
}

/* mara::diff (model_two_body.hpp:492-520) */
static double wrap_delta(double delta, double period)
{
    double a = delta, b = delta + period, c = delta - period;
    if (fabs(a) < fmin(fabs(b), fabs(c))) return a;
    if (fabs(b) < fabs(c)) return b;
    return c;
}

static m3o_elements_t elements_diff(const m3o_elements_t* a, const m3o_elements_t* b)
{
    m3o_elements_t r;
    r.v[0] = wrap_delta(b->v[0] - a->v[0], 2 * M_PI);
    r.v[1] = wrap_delta(b->v[1] - a->v[1], orbital_period(b->v));
    for (int k = 2; k < 10; ++k) r.v[k] = b->v[k] - a->v[k];
    return r;
}

static m3o_elements_t elements_add(m3o_elements_t a, m3o_elements_t b) { for (int k = 0; k < 10; ++k) a.v[k] = a.v[k] + b.v[k]; return a; }
static m3o_elements_t elements_mul(m3o_elements_t a, double s) { for (int k = 0; k < 10; ++k) a.v[k] = a.v[k] * s; return a; }




/* ===========================================================================
 * Point-wise physics
 * ======================================================================== */

/* math_interpolation.hpp:85-94 */
double m3o_plm_gradient(double yl, double y0, double yr, double theta)
{
    double a = (y0 - yl) * theta;
    double b = (yr - yl) * 0.5;
    double c = (yr - y0) * theta;
    double sa = copysign(1.0, a), sb = copysign(1.0, b), sc = copysign(1.0, c);
    double minabs = fmin(fmin(fabs(a), fabs(b)), fabs(c));
    return 0.25 * fabs(sa + sb) * (sa + sc) * minabs;
}

/* primitive_t::to_conserved_angmom_per_area (physics_iso2d.hpp:263-272), the expression create_mesh uses for the initial state */
void m3o_to_conserved_angmom(const double* p, double x, double y, double* q)
{
    q[0] = p[0];
    q[1] = p[0] * (x * p[1] + y * p[2]);
    q[2] = p[0] * (x * p[2] - y * p[1]);
}

/* recover_primitive(Q, x) (physics_iso2d.hpp:376-389), the expression advance_q uses for every cell */
int m3o_recover_primitive_q(const double* q, double x, double y, double* p)
{
    double sigma = q[0];
    double sr = q[1] / sigma;
    double lz = q[2] / sigma;
    double r2 = x * x + y * y;
    p[0] = sigma;
    p[1] = (sr * x - lz * y) / r2;
    p[2] = (sr * y + lz * x) / r2;
    return sigma < 0.0;
}

/* physics_iso2d.hpp:488-506 with flux (:299-307), wavespeeds (:320-328), to_conserved_per_area (:249-258).
 * nhat = on_axis(axis): (1,0,0) or (0,1,0) (core_geometric.hpp:62-74). */
void m3o_riemann_hlle(const double* pl, const double* pr, double cs2, int axis, double* flux)
{
    double n1 = axis == 0 ? 1.0 : 0.0;
    double n2 = axis == 0 ? 0.0 : 1.0;
    double n3 = 0.0;
    double Ul[3] = {pl[0], pl[0] * pl[1], pl[0] * pl[2]};
    double Ur[3] = {pr[0], pr[0] * pr[1], pr[0] * pr[2]};
    double cs = sqrt(cs2);
    double vl = pl[1] * n1 + pl[2] * n2 + 0.0 * n3;
    double vr = pr[1] * n1 + pr[2] * n2 + 0.0 * n3;
    double alm = vl - cs, alp = vl + cs, arm = vr - cs, arp = vr + cs;
    double pgl = pl[0] * cs2, pgr = pr[0] * cs2;
    double Fl[3] = {vl * pl[0], vl * pl[0] * pl[1] + pgl * n1, vl * pl[0] * pl[2] + pgl * n2};
    double Fr[3] = {vr * pr[0], vr * pr[0] * pr[1] + pgr * n1, vr * pr[0] * pr[2] + pgr * n2};
    double ap = fmax(0.0, fmax(alp, arp));
    double am = fmin(0.0, fmin(alm, arm));

    for (int c = 0; c < 3; ++c)
    {
        flux[c] = (Fl[c] * ap - Fr[c] * am - (Ul[c] - Ur[c]) * ap * am) / (ap - am);
    }
}

/* grav_phi_field: scheme.cpp:101-111 */
static double grav_phi(double rs, double bx, double by, double bm, double x, double y)
{
    double G = 1.0;
    double drx = x - bx, dry = y - by;
    double dr2 = drx * drx + dry * dry;
    double rs2 = rs * rs;
    return -G * bm / pow(dr2 + rs2, 0.5);
}

/* cs2_at_position: scheme.cpp:160-175 */
static double cs2_at(const m3o_config_t* c, const m3o_two_body_t* tb, double x, double y)
{
    double M = c->mach_number;
    if (c->axisymmetric_cs2)
    {
        double GM = 1.0;
        double r2 = x * x + y * y;
        return GM / sqrt(r2) / M / M;
    }
    double phi1 = grav_phi(c->softening_radius, tb->b[0][1], tb->b[0][2], tb->b[0][0], x, y);
    double phi2 = grav_phi(c->softening_radius, tb->b[1][1], tb->b[1][2], tb->b[1][0], x, y);
    return -(phi1 + phi2) / M / M;
}

/* nu_at_position: scheme.cpp:177-193 */
static double nu_at(const m3o_config_t* c, double x, double y, double cs2)
{
    double radius = sqrt(x * x + y * y);
    double rc = c->alpha_cutoff_radius;
    double profile = rc > 0.0 ? 0.5 * (1.0 + tanh(3.0 * (radius - rc))) : 1.0;
    if (c->nu > 0.0)
    {
        return profile * c->nu;
    }
    return profile * c->alpha * sqrt(cs2) * (radius / c->mach_number);
}

/* viscous_flux: scheme.cpp:220-262.  gl, gr: longitudinal gradients; hl, hr: transverse. */
static void viscous_flux(int axis, const double* gl, const double* gr, const double* hl, const double* hr, double mu, double* f)
{
    if (axis == 0)
    {
        double dx_ux = 0.5 * (gl[1] + gr[1]);
        double dx_uy = 0.5 * (gl[2] + gr[2]);
        double dy_ux = 0.5 * (hl[1] + hr[1]);
        double dy_uy = 0.5 * (hl[2] + hr[2]);
        double tauxx = mu * (dx_ux - dy_uy);
        double tauxy = mu * (dx_uy + dy_ux);
        f[0] = 0.0; f[1] = -tauxx; f[2] = -tauxy;
    }
    else
    {
        double dx_ux = 0.5 * (hl[1] + hr[1]);
        double dx_uy = 0.5 * (hl[2] + hr[2]);
        double dy_ux = 0.5 * (gl[1] + gr[1]);
        double dy_uy = 0.5 * (gl[2] + gr[2]);
        double tauyx =  mu * (dx_uy + dy_ux);
        double tauyy = -mu * (dx_ux - dy_uy);
        f[0] = 0.0; f[1] = -tauyx; f[2] = -tauyy;
    }
}

/* intercell_flux_u: scheme.cpp:268-293 */
/* to_angmom_fluxes (scheme.cpp:199-214; physics_iso2d.hpp:434-441): F(Sr) = x F(px) + y F(py), F(Lz) = x F(py) - y F(px),
 * and no angular momentum leaves through the domain edge */
static void to_angmom_fluxes(int axis, double rd, double x, double y, double* f)
{
    double flux_px = f[1], flux_py = f[2];
    double flux_sr = x * flux_px + y * flux_py;
    double flux_lz = x * flux_py - y * flux_px;
    if (axis == 0 && (x == -rd || x == rd)) flux_lz = 0.0;
    if (axis == 1 && (y == -rd || y == rd)) flux_lz = 0.0;
    f[1] = flux_sr;
    f[2] = flux_lz;
}

static void intercell_flux_u(const m3o_config_t* c, const m3o_two_body_t* tb, int axis, double grid_spacing,
    double xf, double yf, const double* pl, const double* pr, const double* gl, const double* gr,
    const double* hl, const double* hr, double* flux)
{
    double pl_hat[3], pr_hat[3], fh[3], fv[3];
    for (int k = 0; k < 3; ++k)
    {
        pl_hat[k] = pl[k] + gl[k] * 0.5 * grid_spacing;
        pr_hat[k] = pr[k] - gr[k] * 0.5 * grid_spacing;
    }
    double cs2 = cs2_at(c, tb, xf, yf);
    double nu  = nu_at(c, xf, yf, cs2);
    double mu  = 0.5 * nu * (pl_hat[0] + pr_hat[0]);
    m3o_riemann_hlle(pl_hat, pr_hat, cs2, axis, fh);
    viscous_flux(axis, gl, gr, hl, hr, mu, fv);
    for (int k = 0; k < 3; ++k) flux[k] = fh[k] + fv[k];
}




/* ===========================================================================
 * Guard-zone fill: get_cell_block (mesh_tree_operators.hpp:223-252) on a tree of
 * [3][N][N] blocks.  Returns component c at cell (i, j) of the (possibly
 * manufactured) block at tree index (level, I, J).
 * ======================================================================== */
static double cell_block_value(const m3o_mesh_t* m, const double* field, int level, long I, long J, int i, int j, int c)
{
    int N = m->N;
    size_t bs = (size_t) 3 * N * N;
    node_t* leaf = find_leaf(m->root, level, I, J);

    if (leaf)   /* same level: the leaf itself */
    {
        return field[leaf->leaf_id * bs + (size_t) c * N * N + i * N + j];
    }
    leaf = find_leaf(m->root, level - 1, I / 2, J / 2);

    if (leaf)   /* neighbour is coarser: refine_cells<2> is piecewise constant (mesh_prolong_restrict.hpp:161-196, 324-335) */
    {
        int bx = I % 2, by = J % 2;
        int ci = (bx * N + i) / 2, cj = (by * N + j) / 2;
        return field[leaf->leaf_id * bs + (size_t) c * N * N + ci * N + cj];
    }
    /* neighbour is refined: combine_cells of the 4 children, then coarsen_cells<2>
     * = restrict_cells(0) then restrict_cells(1), each `(h0 + h1) / 2`
     * (mesh_prolong_restrict.hpp:124-132, 262-272, 379-381) */
    node_t* n = find_node(m->root, level, I, J);
    if (! n || is_leaf(n)) { fprintf(stderr, "m3o: get_cell_block failed (tree has over-refined neighbors?)\n"); abort(); }
    double r0[2];
    for (int dj = 0; dj < 2; ++dj)
    {
        double h[2];
        for (int di = 0; di < 2; ++di)
        {
            int fi = 2 * i + di, fj = 2 * j + dj;
            node_t* ch = n->child[(fi >= N) + 2 * (fj >= N)];
            if (! is_leaf(ch)) { fprintf(stderr, "m3o: get_cell_block failed (child is not a leaf)\n"); abort(); }
            h[di] = field[ch->leaf_id * bs + (size_t) c * N * N + (fi % N) * N + (fj % N)];
        }
        r0[dj] = (h[0] + h[1]) / 2;
    }
    return (r0[0] + r0[1]) / 2;
}

/* extend(tree, axis, 1) (scheme.cpp:132-142) for block b: out is [3][N+2][N] (axis 0) or [3][N][N+2] (axis 1) */
static void extend_block(const m3o_mesh_t* m, const double* field, int b, int axis, double* out)
{
    int N = m->N;
    int level = (int) m->index[3 * b];
    long I = m->index[3 * b + 1], J = m->index[3 * b + 2], n = 1L << level;
    size_t bs = (size_t) 3 * N * N;
    long Ip = axis == 0 ? (I + n - 1) % n : I, Jp = axis == 1 ? (J + n - 1) % n : J;   /* prev_on (core_tree.hpp:204) */
    long In = axis == 0 ? (I + n + 1) % n : I, Jn = axis == 1 ? (J + n + 1) % n : J;   /* next_on (core_tree.hpp:203) */

    for (int c = 0; c < 3; ++c)
    {
        const double* C = field + b * bs + (size_t) c * N * N;
        if (axis == 0)
        {
            double* o = out + (size_t) c * (N + 2) * N;
            for (int j = 0; j < N; ++j)
            {
                o[j] = cell_block_value(m, field, level, Ip, Jp, N - 1, j, c);
                for (int i = 0; i < N; ++i) o[(i + 1) * N + j] = C[i * N + j];
                o[(N + 1) * N + j] = cell_block_value(m, field, level, In, Jn, 0, j, c);
            }
        }
        else
        {
            double* o = out + (size_t) c * N * (N + 2);
            for (int i = 0; i < N; ++i)
            {
                o[i * (N + 2)] = cell_block_value(m, field, level, Ip, Jp, i, N - 1, c);
                for (int j = 0; j < N; ++j) o[i * (N + 2) + j + 1] = C[i * N + j];
                o[i * (N + 2) + N + 1] = cell_block_value(m, field, level, In, Jn, i, 0, c);
            }
        }
    }
}




/* ===========================================================================
 * source_term_total_t (scheme.cpp:22-35) as 18 doubles, in declaration order:
 * mass_acc[2] angmom_acc[2] torque[2] px_acc[2] py_acc[2] fx[2] fy[2] work[2] mass_ej angmom_ej
 * ======================================================================== */
enum { T_MASS = 0, T_LACC = 2, T_TORQ = 4, T_PXAC = 6, T_PYAC = 8, T_FX = 10, T_FY = 12, T_WORK = 14, T_MEJ = 16, T_LEJ = 17, T_COUNT = 18 };

/* tree.sum() (core_tree.hpp:502 with core_sequence.hpp:216-224): at every node
 * ((((0 + c0) + c1) + c2) + c3), field-wise (scheme.cpp:65-79). */
static void totals_tree_sum(const node_t* n, const double* per_block, double* out)
{
    if (is_leaf(n))
    {
        memcpy(out, per_block + (size_t) n->leaf_id * T_COUNT, sizeof(double) * T_COUNT);
        return;
    }
    double acc[T_COUNT] = {0};
    for (int c = 0; c < 4; ++c)
    {
        double t[T_COUNT];
        totals_tree_sum(n->child[c], per_block, t);
        for (int k = 0; k < T_COUNT; ++k) acc[k] = acc[k] + t[k];
    }
    memcpy(out, acc, sizeof(acc));
}

/* `work` lambda: scheme.cpp:363-374 */
static double work_on(const double* body, const double* du)
{
    double M0 = body[0], px0 = body[3] * M0, py0 = body[4] * M0;
    double M1 = M0 + du[0], px1 = px0 + du[1], py1 = py0 + du[2];
    return ((px1 * px1 + py1 * py1) / M1 - (px0 * px0 + py0 * py0) / M0) * 0.5;
}




/* ===========================================================================
 * advance_u (scheme.cpp:790-904)
 * ======================================================================== */
int m3o_advance(const m3o_mesh_t* m, const m3o_solution_t* in, double dt, int safe_mode, m3o_solution_t* out)
{
    const m3o_config_t* c = &m->config;
    /* conserve_linear_p = 0: advance_q (scheme.cpp:906-1020); in->U then holds conserved_q = (sigma, Sr, Lz) */
    const int qmode = ! c->conserve_linear_p;

    int N = m->N, B = m->B, V = N + 1;
    size_t bs = (size_t) 3 * N * N;
    double th = safe_mode ? 0.0 : c->plm_theta;
    double spacing_at_root = 2.0 * c->domain_radius / N;
    double rs = c->softening_radius;

    double* p0 = (double*) malloc(sizeof(double) * B * bs);
    double* gx = (double*) malloc(sizeof(double) * B * bs);
    double* gy = (double*) malloc(sizeof(double) * B * bs);
    double* fhx = (double*) malloc(sizeof(double) * (size_t) B * 3 * (N + 1) * N);   /* [B][3][N+1][N] */
    double* fhy = (double*) malloc(sizeof(double) * (size_t) B * 3 * N * (N + 1));   /* [B][3][N][N+1] */
    double* ex = (double*) malloc(sizeof(double) * 3 * (N + 2) * N);
    double* ey = (double*) malloc(sizeof(double) * 3 * (N + 2) * N);
    double* gxex = (double*) malloc(sizeof(double) * 3 * (N + 2) * N);
    double* gyex = (double*) malloc(sizeof(double) * 3 * (N + 2) * N);
    double* gxey = (double*) malloc(sizeof(double) * 3 * (N + 2) * N);
    double* gyey = (double*) malloc(sizeof(double) * 3 * (N + 2) * N);
    double* totals = (double*) calloc((size_t) B * T_COUNT, sizeof(double));

    /* P1 recover_primitive (scheme.cpp:1075-1101; physics_iso2d.hpp:351-362) */
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < N * N; ++k)
        {
            const double* U = in->U + b * bs;
            double sigma = U[k];
            p0[b * bs + k] = sigma;
            if (! qmode)
            {
                p0[b * bs + N * N + k]     = U[N * N + k] / sigma;
                p0[b * bs + 2 * N * N + k] = U[2 * N * N + k] / sigma;
            }
            else    /* recover_primitive(Q, x) (physics_iso2d.hpp:376-389) */
            {
                double x = m->centers[(size_t) b * 2 * N * N + k], y = m->centers[(size_t) b * 2 * N * N + N * N + k];
                double sr = U[N * N + k] / sigma;
                double lz = U[2 * N * N + k] / sigma;
                double r2 = x * x + y * y;
                p0[b * bs + N * N + k]     = (sr * x - lz * y) / r2;
                p0[b * bs + 2 * N * N + k] = (sr * y + lz * x) / r2;
            }
        }

    /* P2 + P3 guard fill of p0, PLM gradients / spacing (scheme.cpp:794-809, 148-154) */
    for (int b = 0; b < B; ++b)
    {
        double spacing = spacing_at_root / (1 << m->index[3 * b]);
        extend_block(m, p0, b, 0, ex);
        extend_block(m, p0, b, 1, ey);
        for (int q = 0; q < 3; ++q)
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j)
                {
                    const double* X = ex + (size_t) q * (N + 2) * N;
                    const double* Y = ey + (size_t) q * N * (N + 2);
                    gx[b * bs + (size_t) q * N * N + i * N + j] = m3o_plm_gradient(X[i * N + j], X[(i + 1) * N + j], X[(i + 2) * N + j], th) / spacing;
                    gy[b * bs + (size_t) q * N * N + i * N + j] = m3o_plm_gradient(Y[i * (N + 2) + j], Y[i * (N + 2) + j + 1], Y[i * (N + 2) + j + 2], th) / spacing;
                }
    }

    /* P5 (scheme.cpp:814) */
    m3o_two_body_t tb;
    m3o_two_body_state(&in->elements, in->time, &tb);

    /* P4 + P6 block_fluxes_u (scheme.cpp:472-516) */
    for (int b = 0; b < B; ++b)
    {
        double grid_spacing = 2.0 * c->domain_radius / N / (1 << m->index[3 * b]);
        const double* xv = m->vertices + (size_t) b * 2 * V * V;
        const double* yv = xv + V * V;
        double* Fx = fhx + (size_t) b * 3 * (N + 1) * N;
        double* Fy = fhy + (size_t) b * 3 * N * (N + 1);

        extend_block(m, p0, b, 0, ex);
        extend_block(m, p0, b, 1, ey);
        extend_block(m, gx, b, 0, gxex);
        extend_block(m, gx, b, 1, gxey);
        extend_block(m, gy, b, 0, gyex);
        extend_block(m, gy, b, 1, gyey);

        for (int f = 0; f <= N; ++f)        /* x faces: (N+1, N) */
            for (int j = 0; j < N; ++j)
            {
                double pl[3], pr[3], gl[3], gr[3], hl[3], hr[3], flux[3];
                for (int q = 0; q < 3; ++q)
                {
                    size_t o = (size_t) q * (N + 2) * N;
                    pl[q] = ex[o + f * N + j];   pr[q] = ex[o + (f + 1) * N + j];
                    gl[q] = gxex[o + f * N + j]; gr[q] = gxex[o + (f + 1) * N + j];
                    hl[q] = gyex[o + f * N + j]; hr[q] = gyex[o + (f + 1) * N + j];
                }
                double xf = (xv[f * V + j] + xv[f * V + j + 1]) * 0.5;   /* midpoint_on_axis(1) */
                double yf = (yv[f * V + j] + yv[f * V + j + 1]) * 0.5;
                double dy = yv[f * V + j + 1] - yv[f * V + j];           /* difference_on_axis(1) of y */
                intercell_flux_u(c, &tb, 0, grid_spacing, xf, yf, pl, pr, gl, gr, hl, hr, flux);
                if (qmode) to_angmom_fluxes(0, c->domain_radius, xf, yf, flux);
                for (int q = 0; q < 3; ++q) Fx[(size_t) q * (N + 1) * N + f * N + j] = flux[q] * dy;
            }
        for (int i = 0; i < N; ++i)         /* y faces: (N, N+1) */
            for (int f = 0; f <= N; ++f)
            {
                double pl[3], pr[3], gl[3], gr[3], hl[3], hr[3], flux[3];
                for (int q = 0; q < 3; ++q)
                {
                    size_t o = (size_t) q * N * (N + 2);
                    pl[q] = ey[o + i * (N + 2) + f];   pr[q] = ey[o + i * (N + 2) + f + 1];
                    gl[q] = gyey[o + i * (N + 2) + f]; gr[q] = gyey[o + i * (N + 2) + f + 1];
                    hl[q] = gxey[o + i * (N + 2) + f]; hr[q] = gxey[o + i * (N + 2) + f + 1];
                }
                double xf = (xv[i * V + f] + xv[(i + 1) * V + f]) * 0.5;   /* midpoint_on_axis(0) */
                double yf = (yv[i * V + f] + yv[(i + 1) * V + f]) * 0.5;
                double dx = xv[(i + 1) * V + f] - xv[i * V + f];           /* difference_on_axis(0) of x */
                intercell_flux_u(c, &tb, 1, grid_spacing, xf, yf, pl, pr, gl, gr, hl, hr, flux);
                if (qmode) to_angmom_fluxes(1, c->domain_radius, xf, yf, flux);
                for (int q = 0; q < 3; ++q) Fy[(size_t) q * N * (N + 1) + i * (N + 1) + f] = flux[q] * dx;
            }
    }

    /* P7 correct_fluxes_{x,y} (scheme.cpp:614-720): a block whose face neighbour is
     * refined takes restrict_extrinsic (pairwise SUM, mesh_prolong_restrict.hpp:134-142)
     * of the two fine neighbours' boundary fluxes.  Only boundary rows of coarse
     * blocks change and fine blocks' rows are never corrected, so in-place is safe. */
    for (int b = 0; b < B; ++b)
    {
        int level = (int) m->index[3 * b];
        long I = m->index[3 * b + 1], J = m->index[3 * b + 2], n = 1L << level;
        double* Fx = fhx + (size_t) b * 3 * (N + 1) * N;
        double* Fy = fhy + (size_t) b * 3 * N * (N + 1);

        for (int side = 0; side < 4; ++side)    /* xl, xr, yl, yr */
        {
            int axis = side / 2, upper = side % 2;
            long In = axis == 0 ? (I + n + (upper ? 1 : -1)) % n : I;
            long Jn = axis == 1 ? (J + n + (upper ? 1 : -1)) % n : J;
            if (find_leaf(m->root, level, In, Jn)) continue;
            if (find_leaf(m->root, level - 1, In / 2, Jn / 2)) continue;
            node_t* nb = find_node(m->root, level, In, Jn);

            for (int k = 0; k < N; ++k)
                for (int q = 0; q < 3; ++q)
                {
                    int f0 = 2 * k, f1 = 2 * k + 1;   /* fine face indexes along the shared edge */
                    double h0, h1;
                    if (axis == 0)
                    {
                        /* xl: children {1,0},{1,1} of prev, their last face row; xr: children {0,0},{0,1} of next, first row */
                        int bx = upper ? 0 : 1, row = upper ? 0 : N;
                        node_t* c0 = nb->child[bx + 2 * (f0 >= N)];
                        node_t* c1 = nb->child[bx + 2 * (f1 >= N)];
                        h0 = fhx[(size_t) c0->leaf_id * 3 * (N + 1) * N + (size_t) q * (N + 1) * N + row * N + f0 % N];
                        h1 = fhx[(size_t) c1->leaf_id * 3 * (N + 1) * N + (size_t) q * (N + 1) * N + row * N + f1 % N];
                        Fx[(size_t) q * (N + 1) * N + (upper ? N : 0) * N + k] = h0 + h1;
                    }
                    else
                    {
                        /* yl: children {0,1},{1,1} of prev, their last face column; yr: children {0,0},{1,0} of next, first column */
                        int by = upper ? 0 : 1, col = upper ? 0 : N;
                        node_t* c0 = nb->child[(f0 >= N) + 2 * by];
                        node_t* c1 = nb->child[(f1 >= N) + 2 * by];
                        h0 = fhy[(size_t) c0->leaf_id * 3 * N * (N + 1) + (size_t) q * N * (N + 1) + (f0 % N) * (N + 1) + col];
                        h1 = fhy[(size_t) c1->leaf_id * 3 * N * (N + 1) + (size_t) q * N * (N + 1) + (f1 % N) * (N + 1) + col];
                        Fy[(size_t) q * N * (N + 1) + k * (N + 1) + (upper ? N : 0)] = h0 + h1;
                    }
                }
        }
    }

    /* P8 block_update_u + source_terms_u (scheme.cpp:568-587, 345-411) */
    for (int b = 0; b < B; ++b)
    {
        const double* U  = in->U + b * bs;
        const double* U0 = m->U0 + b * bs;
        const double* xc = m->centers + (size_t) b * 2 * N * N;
        const double* yc = xc + N * N;
        const double* dA = m->areas + (size_t) b * N * N;
        const double* br = m->buffer_rate + (size_t) b * N * N;
        const double* Fx = fhx + (size_t) b * 3 * (N + 1) * N;
        const double* Fy = fhy + (size_t) b * 3 * N * (N + 1);
        double* U1 = out->U + b * bs;
        double* T = totals + (size_t) b * T_COUNT;
        double sink_sum[2][3] = {{0}};

        if (qmode)
        {
            /* block_update_q + source_terms_q (scheme.cpp:589-608, 417-466) */
            double sr2 = m->gst_suppr_radius * m->gst_suppr_radius;
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j)
                {
                    int k = i * N + j;
                    double x = xc[k], y = yc[k];
                    double q0[3] = {U[k], U[N * N + k], U[2 * N * N + k]};
                    double sigma = q0[0];
                    double fg[2][2], s_grav[2][3], s_sink[2][3], s_buffer[3], s_geom[3], dps[2][2];

                    for (int a = 0; a < 2; ++a)
                    {
                        double bxp = tb.b[a][1], byp = tb.b[a][2], bm = tb.b[a][0], G = 1.0;
                        double drx = x - bxp, dry = y - byp;
                        double dr2 = drx * drx + dry * dry;
                        double rs2 = rs * rs;
                        double p15 = pow(dr2 + rs2, 1.5);
                        fg[a][0] = -drx / p15 * G * bm * sigma;
                        fg[a][1] = -dry / p15 * G * bm * sigma;
                        /* force_to_source_terms_q (:333-339), then * dt */
                        s_grav[a][0] = 0.0 * dt;
                        s_grav[a][1] = (x * fg[a][0] + y * fg[a][1]) * dt;
                        s_grav[a][2] = (x * fg[a][1] - y * fg[a][0]) * dt;

                        double s2 = c->sink_radius * c->sink_radius;
                        double a2 = (drx * drx + dry * dry) / s2 / 2.0;
                        double rate = c->sink_rate * exp(-a2);
                        for (int q = 0; q < 3; ++q) s_sink[a][q] = -q0[q] * rate * dt;

                        /* to_conserved_per_area(s_sink, xc) | momentum_vector (physics_iso2d.hpp:402-417) */
                        double r2 = x * x + y * y;
                        dps[a][0] = (s_sink[a][1] * x - s_sink[a][2] * y) / r2;
                        dps[a][1] = (s_sink[a][1] * y + s_sink[a][2] * x) / r2;
                    }
                    for (int q = 0; q < 3; ++q) s_buffer[q] = (U0[(size_t) q * N * N + k] - q0[q]) * br[k] * dt;
                    {
                        /* source_terms_conserved_angmom (physics_iso2d.hpp:277-285) with the ramp (:439-444) */
                        const double* pp = p0 + b * bs;
                        double vx = pp[N * N + k], vy = pp[2 * N * N + k];
                        double ramp = 1.0 - exp(-(x * x + y * y) / sr2);
                        double cs2 = cs2_at(c, &tb, x, y);
                        double Ek = 0.5 * sigma * (vx * vx + vy * vy);
                        double pg = sigma * cs2;
                        s_geom[0] = 0.0 * ramp * dt;
                        s_geom[1] = (Ek + pg) * 2.0 * ramp * dt;
                        s_geom[2] = 0.0 * ramp * dt;
                    }
                    for (int a = 0; a < 2; ++a)
                    {
                        T[T_MASS + a] = T[T_MASS + a] + s_sink[a][0] * dA[k];
                        T[T_LACC + a] = T[T_LACC + a] + s_sink[a][2] * dA[k];
                        T[T_TORQ + a] = T[T_TORQ + a] + s_grav[a][2] * dA[k];
                        T[T_FX + a]   = T[T_FX + a]   + fg[a][0] * dt * dA[k];
                        T[T_FY + a]   = T[T_FY + a]   + fg[a][1] * dt * dA[k];
                        T[T_PXAC + a] = T[T_PXAC + a] + dps[a][0] * dA[k];
                        T[T_PYAC + a] = T[T_PYAC + a] + dps[a][1] * dA[k];
                    }
                    T[T_LEJ] = T[T_LEJ] + s_buffer[2] * dA[k];
                    T[T_MEJ] = T[T_MEJ] + s_buffer[0] * dA[k];

                    for (int q = 0; q < 3; ++q)
                    {
                        double lx = Fx[(size_t) q * (N + 1) * N + (i + 1) * N + j] - Fx[(size_t) q * (N + 1) * N + i * N + j];
                        double ly = Fy[(size_t) q * N * (N + 1) + i * (N + 1) + j + 1] - Fy[(size_t) q * N * (N + 1) + i * (N + 1) + j];
                        double s = s_grav[0][q] + s_grav[1][q] + s_sink[0][q] + s_sink[1][q] + s_buffer[q] + s_geom[q];
                        U1[(size_t) q * N * N + k] = q0[q] - (lx + ly) * dt / dA[k] + s;
                    }
                }
            for (int k = 0; k < T_COUNT; ++k) T[k] = -T[k];
            T[T_WORK] = T[T_WORK + 1] = 0.0;        /* source_terms_q does not set work_done_on */
            continue;
        }
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j)
            {
                int k = i * N + j;
                double x = xc[k], y = yc[k];
                double u0[3] = {U[k], U[N * N + k], U[2 * N * N + k]};
                double sigma = u0[0];
                double fg[2][2], s_grav[2][3], s_sink[2][3], s_buffer[3], s_floor[3];

                for (int a = 0; a < 2; ++a)
                {
                    /* grav_vdot_field (scheme.cpp:85-95) * sigma (:377-378) */
                    double bxp = tb.b[a][1], byp = tb.b[a][2], bm = tb.b[a][0], G = 1.0;
                    double drx = x - bxp, dry = y - byp;
                    double dr2 = drx * drx + dry * dry;
                    double rs2 = rs * rs;
                    double p15 = pow(dr2 + rs2, 1.5);
                    fg[a][0] = -drx / p15 * G * bm * sigma;
                    fg[a][1] = -dry / p15 * G * bm * sigma;
                    s_grav[a][0] = 0.0 * dt;                 /* force_to_source_terms_u (:327-331), then * dt (:380-381) */
                    s_grav[a][1] = fg[a][0] * dt;
                    s_grav[a][2] = fg[a][1] * dt;

                    /* sink_rate_field (scheme.cpp:117-126); s_sink = -u0 * rate * dt (:382-383) */
                    double s2 = c->sink_radius * c->sink_radius;
                    double a2 = (drx * drx + dry * dry) / s2 / 2.0;
                    double rate = c->sink_rate * exp(-a2);
                    for (int q = 0; q < 3; ++q) s_sink[a][q] = -u0[q] * rate * dt;
                }
                for (int q = 0; q < 3; ++q)
                {
                    s_buffer[q] = (U0[(size_t) q * N * N + k] - u0[q]) * br[k] * dt;           /* :384 */
                    s_floor[q]  = u0[q] * 1e-2 * (double) (u0[0] < m->density_floor);          /* :385-388 */
                }
                /* totals: sequential row-major sums starting from zero (:390-408; core_ndarray.hpp:1882-1898) */
                for (int a = 0; a < 2; ++a)
                {
                    T[T_MASS + a] = T[T_MASS + a] + s_sink[a][0] * dA[k];
                    T[T_LACC + a] = T[T_LACC + a] + (x * s_sink[a][2] - y * s_sink[a][1]) * dA[k];
                    T[T_TORQ + a] = T[T_TORQ + a] + (x * s_grav[a][2] - y * s_grav[a][1]) * dA[k];
                    T[T_FX + a]   = T[T_FX + a]   + fg[a][0] * dt * dA[k];
                    T[T_FY + a]   = T[T_FY + a]   + fg[a][1] * dt * dA[k];
                    T[T_PXAC + a] = T[T_PXAC + a] + s_sink[a][1] * dA[k];
                    T[T_PYAC + a] = T[T_PYAC + a] + s_sink[a][2] * dA[k];
                    for (int q = 0; q < 3; ++q) sink_sum[a][q] = sink_sum[a][q] + s_sink[a][q] * dA[k];
                }
                T[T_LEJ] = T[T_LEJ] + (x * s_buffer[2] - y * s_buffer[1]) * dA[k];
                T[T_MEJ] = T[T_MEJ] + s_buffer[0] * dA[k];

                for (int q = 0; q < 3; ++q)
                {
                    double lx = Fx[(size_t) q * (N + 1) * N + (i + 1) * N + j] - Fx[(size_t) q * (N + 1) * N + i * N + j];
                    double ly = Fy[(size_t) q * N * (N + 1) + i * (N + 1) + j + 1] - Fy[(size_t) q * N * (N + 1) + i * (N + 1) + j];
                    double s = s_grav[0][q] + s_grav[1][q] + s_sink[0][q] + s_sink[1][q] + s_buffer[q] + s_floor[q];  /* :410 */
                    U1[(size_t) q * N * N + k] = u0[q] - (lx + ly) * dt / dA[k] + s;                                   /* :585 */
                }
            }
        /* every total is the NEGATED sum (:391-406) */
        for (int k = 0; k < T_COUNT; ++k) if (k != T_WORK && k != T_WORK + 1) T[k] = -T[k];
        for (int a = 0; a < 2; ++a)
        {
            double du[3] = {-sink_sum[a][0], -sink_sum[a][1], -sink_sum[a][2]};
            T[T_WORK + a] = work_on(tb.b[a], du);   /* :407-408 */
        }
    }

    /* P9 tree-sum of the per-block totals (scheme.cpp:829-830) */
    double tot[T_COUNT];
    totals_tree_sum(m->root, totals, tot);

    /* P10 orbital-element bookkeeping (scheme.cpp:832-885) */
    double M1 = tb.b[0][0], M2 = tb.b[1][0];
    double px1 = M1 * tb.b[0][3], py1 = M1 * tb.b[0][4], px2 = M2 * tb.b[1][3], py2 = M2 * tb.b[1][4];
    double dM1 = tot[T_MASS], dM2 = tot[T_MASS + 1];
    double dpx1 = tot[T_PXAC], dpy1 = tot[T_PYAC], dpx2 = tot[T_PXAC + 1], dpy2 = tot[T_PYAC + 1];
    double vx1 = (px1 + dpx1) / (M1 + dM1), vy1 = (py1 + dpy1) / (M1 + dM1);
    double vx2 = (px2 + dpx2) / (M2 + dM2), vy2 = (py2 + dpy2) / (M2 + dM2);
    m3o_two_body_t acc = tb, grv = tb;
    acc.b[0][0] = tb.b[0][0] + dM1;
    acc.b[1][0] = tb.b[1][0] + dM2;
    if (! c->no_accretion_force)
    {
        acc.b[0][3] = vx1; acc.b[0][4] = vy1; acc.b[1][3] = vx2; acc.b[1][4] = vy2;
    }
    grv.b[0][3] = tb.b[0][3] + tot[T_FX]     / tb.b[0][0];
    grv.b[0][4] = tb.b[0][4] + tot[T_FY]     / tb.b[0][0];
    grv.b[1][3] = tb.b[1][3] + tot[T_FX + 1] / tb.b[1][0];
    grv.b[1][4] = tb.b[1][4] + tot[T_FY + 1] / tb.b[1][0];

    double live = in->time > c->begin_live_binary ? 1.0 : 0.0;
    m3o_elements_t E0 = in->elements, E_acc, E_grv;
    int unbound = m3o_orbital_elements(&acc, in->time, &E_acc) | m3o_orbital_elements(&grv, in->time, &E_grv);
    if (unbound) { fprintf(stderr, "m3o: two_body_state does not correspond to a bound orbit\n"); abort(); }

    m3o_elements_t d_acc = elements_diff(&E0, &E_acc);
    m3o_elements_t d_grv = elements_diff(&E0, &E_grv);
    m3o_elements_t d_cm  = elements_zeros();       /* diff_cm: model_two_body.hpp:522-529 */
    d_cm.v[2] = E0.v[4] * dt;
    d_cm.v[3] = E0.v[5] * dt;

    /* the updated state (scheme.cpp:889-903) */
    out->B = in->B; out->N = in->N;
    out->time = in->time + dt;
    out->iter_num = in->iter_num + in->iter_den;   /* iteration + 1 for a rational */
    out->iter_den = in->iter_den;
    for (int a = 0; a < 2; ++a)
    {
        out->mass_accreted_on[a]             = in->mass_accreted_on[a]             + tot[T_MASS + a];
        out->angular_momentum_accreted_on[a] = in->angular_momentum_accreted_on[a] + tot[T_LACC + a];
        out->integrated_torque_on[a]         = in->integrated_torque_on[a]         + tot[T_TORQ + a];
        out->work_done_on[a]                 = in->work_done_on[a]                 + tot[T_WORK + a];
    }
    out->mass_ejected             = in->mass_ejected             + tot[T_MEJ];
    out->angular_momentum_ejected = in->angular_momentum_ejected + tot[T_LEJ];
    out->elements_acc  = elements_add(in->elements_acc, d_acc);
    out->elements_grav = elements_add(in->elements_grav, d_grv);
    out->elements      = elements_add(in->elements, elements_mul(elements_add(elements_add(d_acc, d_grv), d_cm), live));

    /* P11 validate_u (scheme.cpp:726-752) */
    int any_failures = 0;
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < N * N; ++k)
            if (out->U[b * bs + k] < 0.0)
            {
                printf("negative density %3.2e (at position [%+3.2lf %+3.2lf])\n", out->U[b * bs + k],
                    m->centers[(size_t) b * 2 * N * N + k], m->centers[(size_t) b * 2 * N * N + N * N + k]);
                any_failures = 1;
            }

    free(p0); free(gx); free(gy); free(fhx); free(fhy); free(ex); free(ey);
    free(gxex); free(gyex); free(gxey); free(gyey); free(totals);
    return any_failures;
}




/* ===========================================================================
 * maximum_timestep (scheme.cpp:1107-1126) with max_wavespeed (physics_iso2d.hpp:330-337)
 * ======================================================================== */
double m3o_maximum_timestep(const m3o_mesh_t* m, const m3o_solution_t* s)
{
    const m3o_config_t* c = &m->config;
    int N = m->N, B = m->B;
    size_t bs = (size_t) 3 * N * N;
    double spacing_at_root = 2.0 * c->domain_radius / N;
    double result = INFINITY;
    m3o_two_body_t tb;
    m3o_two_body_state(&s->elements, s->time, &tb);

    for (int b = 0; b < B; ++b)
    {
        const double* U = s->U + b * bs;
        const double* xc = m->centers + (size_t) b * 2 * N * N;
        const double* yc = xc + N * N;
        double spacing = spacing_at_root / (1 << m->index[3 * b]);
        double max_wavespeed = -INFINITY;

        for (int k = 0; k < N * N; ++k)
        {
            double sigma = U[k], vx = U[N * N + k] / sigma, vy = U[2 * N * N + k] / sigma;
            if (! c->conserve_linear_p)     /* (the reference itself cannot take this path: it needs fixed_dt = 1 with conserved_q) */
            {
                double sr = vx, lz = vy, r2 = xc[k] * xc[k] + yc[k] * yc[k];
                vx = (sr * xc[k] - lz * yc[k]) / r2;
                vy = (sr * yc[k] + lz * xc[k]) / r2;
            }
            double cs = sqrt(cs2_at(c, &tb, xc[k], yc[k]));
            double ax = fmax(fabs(vx - cs), fabs(vx + cs));
            double ay = fmax(fabs(vy - cs), fabs(vy + cs));
            double a = fmax(ax, ay);
            if (a > max_wavespeed) max_wavespeed = a;
        }
        double dt = spacing / max_wavespeed;
        if (dt < result) result = dt;
    }
    return result;
}




/* ===========================================================================
 * next_solution (subprog_binary.cpp:258-293) and the RK combination
 * solution_t::operator+ / operator* (scheme.cpp:1033-1069)
 * ======================================================================== */
static int igcd(int a, int b) { a = abs(a); b = abs(b); while (b) { int t = a % b; a = b; b = t; } return a ? a : 1; }

static void rk_combine(const m3o_solution_t* s0, const m3o_solution_t* s2, double b0, m3o_solution_t* r)
{
    /* s0 * b0 + s2 * (1 - b0) with b0 = 1/2; every field is a*b0 + b*(1-b0) */
    double b1 = 1.0 - b0;
    size_t n = (size_t) s0->B * 3 * s0->N * s0->N;
    for (size_t k = 0; k < n; ++k) r->U[k] = s0->U[k] * b0 + s2->U[k] * b1;
    r->time = s0->time * b0 + s2->time * b1;
    /* rational: n0/d0 * 1/2 + n2/d2 * 1/2, reduced (core_rational.hpp) */
    int num = s0->iter_num * 2 * s2->iter_den + s2->iter_num * 2 * s0->iter_den;
    int den = 2 * s0->iter_den * 2 * s2->iter_den;
    int g = igcd(num, den);
    r->iter_num = num / g; r->iter_den = den / g;
    for (int a = 0; a < 2; ++a)
    {
        r->mass_accreted_on[a]             = s0->mass_accreted_on[a] * b0             + s2->mass_accreted_on[a] * b1;
        r->angular_momentum_accreted_on[a] = s0->angular_momentum_accreted_on[a] * b0 + s2->angular_momentum_accreted_on[a] * b1;
        r->integrated_torque_on[a]         = s0->integrated_torque_on[a] * b0         + s2->integrated_torque_on[a] * b1;
        r->work_done_on[a]                 = s0->work_done_on[a] * b0                 + s2->work_done_on[a] * b1;
    }
    r->mass_ejected             = s0->mass_ejected * b0             + s2->mass_ejected * b1;
    r->angular_momentum_ejected = s0->angular_momentum_ejected * b0 + s2->angular_momentum_ejected * b1;
    r->elements_acc  = elements_add(elements_mul(s0->elements_acc, b0),  elements_mul(s2->elements_acc, b1));
    r->elements_grav = elements_add(elements_mul(s0->elements_grav, b0), elements_mul(s2->elements_grav, b1));
    r->elements      = elements_add(elements_mul(s0->elements, b0),      elements_mul(s2->elements, b1));
}

static int can_fail(const m3o_mesh_t* m, m3o_solution_t* s, double dt, int safe_mode)
{
    m3o_solution_t* s1 = m3o_solution_clone(s);
    int failed = 0;

    if (m->config.rk_order == 1)
    {
        failed = m3o_advance(m, s, dt, safe_mode, s1);
        if (! failed) solution_assign(s, s1);
    }
    else
    {
        m3o_solution_t* s2 = m3o_solution_clone(s);
        failed = m3o_advance(m, s, dt, safe_mode, s1);
        if (! failed) failed = m3o_advance(m, s1, dt, safe_mode, s2);
        if (! failed)
        {
            rk_combine(s, s2, 0.5, s1);
            s1->B = s->B; s1->N = s->N;
            solution_assign(s, s1);
        }
        m3o_solution_destroy(s2);
    }
    m3o_solution_destroy(s1);
    return failed;
}

int m3o_next_solution(const m3o_mesh_t* m, m3o_solution_t* s, double* dt_used)
{
    double dt = m->config.fixed_dt ? m->recommended_time_step : m->config.cfl_number * m3o_maximum_timestep(m, s);

    if (! can_fail(m, s, dt, 0))
    {
        if (dt_used) *dt_used = dt;
        return 0;
    }
    printf("negative density in updated state\n");
    if (can_fail(m, s, dt * 0.1, 1))
    {
        fprintf(stderr, "m3o: negative density in safe mode\n");
        abort();
    }
    if (dt_used) *dt_used = dt * 0.1;
    return 1;
}
