/**
 * kernels.cu -- CUDA kernels (sm_100a) and launch logic of the iso2d `binary` update.
 *
 * Replaces, per RK stage, the reference's phases P1-P8 and P11 of binary::advance_u
 * (Mara3 src/subprog_binary_scheme.cpp:790-904) and binary::maximum_timestep
 * (:1107-1126).  Two paths produce the same update:
 *
 *   stage_fused<TX,TY>   for "regular" blocks (all 8 neighbours are same-level leaves):
 *                        one CTA per TX x TY tile; the tile plus a 2-cell halo is read
 *                        once from HBM, primitives / PLM differences / face fluxes live
 *                        in shared memory, and the updated cells are written once.
 *   general_*            for blocks touching a refinement jump: guard values are
 *                        fetched through per-face neighbour tables with the reference's
 *                        prolongation (injection) / restriction (2x2 mean) rules for
 *                        primitives AND gradients (mesh_tree_operators.hpp:223-252), and
 *                        coarse faces next to finer blocks take the sum of the two fine
 *                        fluxes (scheme.cpp:614-720).
 *
 * Each CTA also reduces the 16 source-term sums, the CFL minimum and the
 * negative-density count; finish_stage folds the per-CTA rows in a fixed order.
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "comm.hpp"
#include "device_solver.hpp"
#include "iso2d_device.cuh"
#include "kernel_common.cuh"
#include "device_solver_impl.cuh"

using namespace m3b;
using namespace m3b::dev;


namespace
{
    // =======================================================================
    // Block-wide reduction of the stage outputs
    // =======================================================================
    __device__ __forceinline__ double warp_sum(double v)
    {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }

    __device__ __forceinline__ double warp_min(double v)
    {
        for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }

    /**
     * Fold the per-thread sums / CFL minimum into one row.  `red` is shared scratch of
     * (THREADS / 32) * (NUM_SUMS + 1) doubles.  `mask` says which groups of sums can be
     * non-zero anywhere in the CTA (bit 0: gravity, bit 1: sinks, bit 2: buffer).
     */
    __device__ void reduce_and_store(double* red, const double sums[NUM_SUMS], double dtmin, double scale, double* row)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = THREADS / 32;

        {
            // sixteen sums over 32 lanes by recursive halving (a lane keeps half of its values and trades the other half at
            // distances 16, 8, 4, 2; one butterfly step finishes): 16 shuffled doubles per lane instead of 80.
            // Lane 2 j ends with the total of sum j.
            static_assert(NUM_SUMS == 16, "the halving below is written for sixteen values");
            double v8[8], v4[4], v2[2], v1;
            const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
            #pragma unroll
            for (int k = 0; k < 8; ++k) v8[k] = (b16 ? sums[8 + k] : sums[k]) + __shfl_xor_sync(0xffffffffu, b16 ? sums[k] : sums[8 + k], 16);
            #pragma unroll
            for (int k = 0; k < 4; ++k) v4[k] = (b8 ? v8[4 + k] : v8[k]) + __shfl_xor_sync(0xffffffffu, b8 ? v8[k] : v8[4 + k], 8);
            #pragma unroll
            for (int k = 0; k < 2; ++k) v2[k] = (b4 ? v4[2 + k] : v4[k]) + __shfl_xor_sync(0xffffffffu, b4 ? v4[k] : v4[2 + k], 4);
            v1 = (b2 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? v2[0] : v2[1], 2);
            v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
            if ((lane & 1) == 0) red[warp * (NUM_SUMS + 1) + (lane >> 1)] = v1;
        }
        double m = warp_min(dtmin);
        if (lane == 0) red[warp * (NUM_SUMS + 1) + NUM_SUMS] = m;
        __syncthreads();

        if (threadIdx.x < NUM_SUMS)
        {
            double v = 0.0;
            for (int w = 0; w < nw; ++w) v += red[w * (NUM_SUMS + 1) + threadIdx.x];
            row[threadIdx.x] = v * scale;
        }
        if (threadIdx.x == NUM_SUMS)
        {
            double v = red[NUM_SUMS];
            for (int w = 1; w < nw; ++w) v = fmin(v, red[w * (NUM_SUMS + 1) + NUM_SUMS]);
            row[NUM_SUMS] = v;
        }
    }



    // =======================================================================
    // Fused stage kernel for regular blocks
    // =======================================================================
    template<int TX, int TY>
    struct tile_t
    {
        static constexpr int PX = TX + 4, PY = TY + 4;      // primitives: tile + 2 halo
        static constexpr int GX = TX + 2, GY = TY + 2;      // PLM differences: tile + 1 halo
        double P[3][PX][PY];
        double G[6][GX][GY];                                // d/dx (s, vx, vy), d/dy (s, vx, vy), un-divided
        double Fx[3][TX + 1][TY];
        double Fy[3][TX][TY + 1];
        double xv[TX + 1];
        double yv[TY + 1];
        double red[(THREADS / 32) * (NUM_SUMS + 1)];
    };

    template<int TX, int TY>
    __global__ void __launch_bounds__(THREADS, 2) stage_fused(
        mesh_dev_t mesh, model_t model, const stage_t* __restrict__ stage_ptr, const int* __restrict__ regular_list,
        const double* __restrict__ Uin, const double* __restrict__ Un, double* __restrict__ Uout,
        double* partials, fail_dev_t* fail)
    {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        tile_t<TX, TY>& T = *reinterpret_cast<tile_t<TX, TY>*>(smem_raw);

        const stage_t S = *stage_ptr;
        const int N = mesh.N;
        const int tiles_y = N / TY, tiles_per_block = (N / TX) * tiles_y;
        const int b  = regular_list[blockIdx.x / tiles_per_block];
        const int t  = blockIdx.x % tiles_per_block;
        const int i0 = (t / tiles_y) * TX, j0 = (t % tiles_y) * TY;
        const size_t FS = mesh.FS;
        const int tid = threadIdx.x;

        // ---- phase 0: tile + 2-cell halo -> primitives in shared memory (P1 + P2 of advance_u)
        for (int k = tid; k < T.PX * T.PY; k += THREADS)
        {
            int li = k / T.PY, lj = k % T.PY;
            int gi = i0 - 2 + li, gj = j0 - 2 + lj;
            int di = gi < 0 ? -1 : (gi >= N ? 1 : 0);
            int dj = gj < 0 ? -1 : (gj >= N ? 1 : 0);
            int nb = (di | dj) ? mesh.nbr9[b * 9 + (di + 1) * 3 + (dj + 1)] : b;
            size_t c = (size_t(nb) * N + (gi - di * N)) * N + (gj - dj * N);
            prim_t p = cons_to_prim(Uin[c], Uin[FS + c], Uin[2 * FS + c]);
            T.P[0][li][lj] = p.s;
            T.P[1][li][lj] = p.vx;
            T.P[2][li][lj] = p.vy;
        }
        if (tid <= TX) T.xv[tid] = mesh.xv[size_t(b) * (N + 1) + i0 + tid];
        if (tid >= 64 && tid - 64 <= TY) T.yv[tid - 64] = mesh.yv[size_t(b) * (N + 1) + j0 + tid - 64];
        __syncthreads();

        // ---- phase 1: PLM differences on tile + 1 halo (P3; the guard gradients of P4 are the neighbours' own)
        for (int k = tid; k < T.GX * T.GY; k += THREADS)
        {
            int li = k / T.GY, lj = k % T.GY;       // gradient cell (li, lj) <-> primitive cell (li + 1, lj + 1)
            #pragma unroll
            for (int q = 0; q < 3; ++q)
            {
                double c = T.P[q][li + 1][lj + 1];
                T.G[q][li][lj]     = plm_diff(T.P[q][li][lj + 1], c, T.P[q][li + 2][lj + 1], S.theta);
                T.G[3 + q][li][lj] = plm_diff(T.P[q][li + 1][lj], c, T.P[q][li + 1][lj + 2], S.theta);
            }
        }
        __syncthreads();

        // ---- phase 2: HLLE + viscous fluxes on the (TX+1) x TY x-faces and TX x (TY+1) y-faces (P6)
        const double h = mesh.spacing[b], inv_h = 1.0 / h;

        auto x_face = [&] (int li, int lj)      // face between tile cells (li - 1, lj) and (li, lj)
        {
            eos_t e = eos_at_face(model, S, T.xv[li], 0.5 * (T.yv[lj] + T.yv[lj + 1]));
            prim_t pl = {T.P[0][li + 1][lj + 2], T.P[1][li + 1][lj + 2], T.P[2][li + 1][lj + 2]};
            prim_t pr = {T.P[0][li + 2][lj + 2], T.P[1][li + 2][lj + 2], T.P[2][li + 2][lj + 2]};
            prim_t gl = {T.G[0][li][lj + 1], T.G[1][li][lj + 1], T.G[2][li][lj + 1]};
            prim_t gr = {T.G[0][li + 1][lj + 1], T.G[1][li + 1][lj + 1], T.G[2][li + 1][lj + 1]};
            double F[3];
            face_flux<0>(e, pl, pr, gl, gr, T.G[4][li][lj + 1], T.G[5][li][lj + 1], T.G[4][li + 1][lj + 1], T.G[5][li + 1][lj + 1], 0.5, inv_h, F);
            T.Fx[0][li][lj] = F[0]; T.Fx[1][li][lj] = F[1]; T.Fx[2][li][lj] = F[2];
        };
        auto y_face = [&] (int li, int lj)      // face between tile cells (li, lj - 1) and (li, lj)
        {
            eos_t e = eos_at_face(model, S, 0.5 * (T.xv[li] + T.xv[li + 1]), T.yv[lj]);
            prim_t pl = {T.P[0][li + 2][lj + 1], T.P[1][li + 2][lj + 1], T.P[2][li + 2][lj + 1]};
            prim_t pr = {T.P[0][li + 2][lj + 2], T.P[1][li + 2][lj + 2], T.P[2][li + 2][lj + 2]};
            prim_t gl = {T.G[3][li + 1][lj], T.G[4][li + 1][lj], T.G[5][li + 1][lj]};
            prim_t gr = {T.G[3][li + 1][lj + 1], T.G[4][li + 1][lj + 1], T.G[5][li + 1][lj + 1]};
            double F[3];
            face_flux<1>(e, pl, pr, gl, gr, T.G[1][li + 1][lj], T.G[2][li + 1][lj], T.G[1][li + 1][lj + 1], T.G[2][li + 1][lj + 1], 0.5, inv_h, F);
            T.Fy[0][li][lj] = F[0]; T.Fy[1][li][lj] = F[1]; T.Fy[2][li][lj] = F[2];
        };
        for (int k = tid; k < TX * TY; k += THREADS)
        {
            x_face(k / TY, k % TY);
            y_face(k / TY, k % TY);
        }
        for (int k = tid; k < TX + TY; k += THREADS)
        {
            if (k < TY) x_face(TX, k); else y_face(k - TY, TY);
        }
        __syncthreads();

        // ---- phase 3: conservative update + source terms (P8), validation (P11), CFL estimate
        double sums[NUM_SUMS];
        #pragma unroll
        for (int k = 0; k < NUM_SUMS; ++k) sums[k] = 0.0;
        double dtmin = 1e300;
        const double dt_over_h = S.dt * inv_h;

        for (int k = tid; k < TX * TY; k += THREADS)
        {
            int li = k / TY, lj = k % TY;
            size_t c = (size_t(b) * N + (i0 + li)) * N + (j0 + lj);
            double s = Uin[c], px = Uin[FS + c], py = Uin[2 * FS + c];
            double br = mesh.br[c];
            double u0s = 0.0, u0x = 0.0, u0y = 0.0;
            if (br != 0.0) { u0s = mesh.U0[c]; u0x = mesh.U0[FS + c]; u0y = mesh.U0[2 * FS + c]; }

            double x = 0.5 * (T.xv[li] + T.xv[li + 1]), y = 0.5 * (T.yv[lj] + T.yv[lj + 1]);
            double src[3], y1, y2;
            source_terms(model, S, x, y, s, px, py, u0s, u0x, u0y, br, src, sums, y1, y2);

            double n0 = s  - ((T.Fx[0][li + 1][lj] - T.Fx[0][li][lj]) + (T.Fy[0][li][lj + 1] - T.Fy[0][li][lj])) * dt_over_h + src[0];
            double n1 = px - ((T.Fx[1][li + 1][lj] - T.Fx[1][li][lj]) + (T.Fy[1][li][lj + 1] - T.Fy[1][li][lj])) * dt_over_h + src[1];
            double n2 = py - ((T.Fx[2][li + 1][lj] - T.Fx[2][li][lj]) + (T.Fy[2][li][lj + 1] - T.Fy[2][li][lj])) * dt_over_h + src[2];

            if (n0 < 0.0) report_negative(fail, b, (i0 + li) * N + j0 + lj, n0);

            if (S.combine)
            {
                double w = 1.0 - S.rk_b0;
                n0 = Un[c] * S.rk_b0 + n0 * w;
                n1 = Un[FS + c] * S.rk_b0 + n1 * w;
                n2 = Un[2 * FS + c] * S.rk_b0 + n2 * w;
            }
            Uout[c] = n0; Uout[FS + c] = n1; Uout[2 * FS + c] = n2;

            if (S.compute_dt)
            {
                double mx = n1, my = n2;
                if (mesh.qmode) angmom_to_linear(x, y, n1, n2, mx, my);
                dtmin = fmin(dtmin, h / max_wavespeed(model, S, x, y, y1, y2, n0, mx, my));
            }
        }
        reduce_and_store(T.red, sums, dtmin, h * h, partials + size_t(blockIdx.x) * ROW);
    }


}
namespace
{
    // =======================================================================
    // General path: any 2:1 balanced tree
    // =======================================================================

    /** Location of a (possibly guard) cell of block b: which leaves hold it and how to combine them. */
    struct cell_ref_t
    {
        int kind;           // 0: one cell of `leaf[0]`; 2: mean of 2x2 cells spread over up to 4 leaves
        int leaf[4];
        int ci[4], cj[4];
    };

    /** get_cell_block (mesh_tree_operators.hpp:223-252) resolved for one cell (i, j), -1 <= i, j <= N. */
    __device__ __forceinline__ cell_ref_t resolve_cell(const mesh_dev_t& m, int b, int i, int j)
    {
        const int N = m.N;
        cell_ref_t r;
        r.kind = 0;

        if (i >= 0 && i < N && j >= 0 && j < N)
        {
            r.leaf[0] = b; r.ci[0] = i; r.cj[0] = j;
            return r;
        }
        int side = i < 0 ? 0 : (i >= N ? 1 : (j < 0 ? 2 : 3));
        int ii = i < 0 ? N - 1 : (i >= N ? 0 : i);
        int jj = j < 0 ? N - 1 : (j >= N ? 0 : j);
        const face_nbr_dev_t nb = m.nbr[b * 4 + side];

        if (nb.kind == 0)           // same level: the neighbour's own cell
        {
            r.leaf[0] = nb.leaf[0]; r.ci[0] = ii; r.cj[0] = jj;
        }
        else if (nb.kind == 1)      // coarser: piecewise-constant prolongation (mesh_prolong_restrict.hpp:161-196)
        {
            r.leaf[0] = nb.leaf[0]; r.ci[0] = (nb.bx * N + ii) / 2; r.cj[0] = (nb.by * N + jj) / 2;
        }
        else                        // finer: 2x2 mean over the children (mesh_prolong_restrict.hpp:124-132, 262-272)
        {
            r.kind = 2;
            #pragma unroll
            for (int q = 0; q < 4; ++q)
            {
                int fi = 2 * ii + (q & 1), fj = 2 * jj + (q >> 1);
                r.leaf[q] = nb.leaf[(fi >= N) + 2 * (fj >= N)];
                r.ci[q] = fi % N; r.cj[q] = fj % N;
            }
        }
        return r;
    }

    __device__ __forceinline__ prim_t load_prim(const mesh_dev_t& m, const double* U, int leaf, int i, int j)
    {
        size_t c = (size_t(leaf) * m.N + i) * m.N + j;
        if (m.qmode)
        {
            // recover_primitive(Q, x) (physics_iso2d.hpp:376-389) at the centre of the cell in ITS block
            const double* xv = m.xv + size_t(leaf) * (m.N + 1);
            const double* yv = m.yv + size_t(leaf) * (m.N + 1);
            const double x = 0.5 * (xv[i] + xv[i + 1]), y = 0.5 * (yv[j] + yv[j + 1]);
            const double s = U[c], sr = U[m.FS + c] / s, lz = U[2 * m.FS + c] / s;
            double vx, vy;
            angmom_to_linear(x, y, sr, lz, vx, vy);
            return {s, vx, vy};
        }
        return cons_to_prim(U[c], U[m.FS + c], U[2 * m.FS + c]);
    }

    /** Primitive at cell (i, j) of block b with guard fill: extend(p0, axis, 1) (scheme.cpp:132-142). */
    __device__ __forceinline__ prim_t prim_from_ref(const mesh_dev_t& m, const double* U, const cell_ref_t& r)
    {
        if (r.kind == 0) return load_prim(m, U, r.leaf[0], r.ci[0], r.cj[0]);
        prim_t p00 = load_prim(m, U, r.leaf[0], r.ci[0], r.cj[0]);
        prim_t p10 = load_prim(m, U, r.leaf[1], r.ci[1], r.cj[1]);
        prim_t p01 = load_prim(m, U, r.leaf[2], r.ci[2], r.cj[2]);
        prim_t p11 = load_prim(m, U, r.leaf[3], r.ci[3], r.cj[3]);
        // restrict on axis 0 then on axis 1, each (h0 + h1) / 2
        return {((p00.s + p10.s) * 0.5 + (p01.s + p11.s) * 0.5) * 0.5,
                ((p00.vx + p10.vx) * 0.5 + (p01.vx + p11.vx) * 0.5) * 0.5,
                ((p00.vy + p10.vy) * 0.5 + (p01.vy + p11.vy) * 0.5) * 0.5};
    }

    __device__ __forceinline__ prim_t prim_at(const mesh_dev_t& m, const double* U, int b, int i, int j)
    {
        return prim_from_ref(m, U, resolve_cell(m, b, i, j));
    }

    __device__ __forceinline__ prim_t load_grad(const mesh_dev_t& m, const double* G, int axis, int leaf, int i, int j)
    {
        size_t c = (size_t(m.gslot[leaf]) * m.N + i) * m.N + j;
        const double* g = G + size_t(3 * axis) * m.GS;
        return {g[c], g[m.GS + c], g[2 * m.GS + c]};
    }

    /** Gradient (d/d axis) at cell (i, j) of block b with guard fill: extend(gx, ...) etc. (scheme.cpp:810-813). */
    __device__ __forceinline__ prim_t grad_from_ref(const mesh_dev_t& m, const double* G, int axis, const cell_ref_t& r)
    {
        if (r.kind == 0) return load_grad(m, G, axis, r.leaf[0], r.ci[0], r.cj[0]);
        prim_t g00 = load_grad(m, G, axis, r.leaf[0], r.ci[0], r.cj[0]);
        prim_t g10 = load_grad(m, G, axis, r.leaf[1], r.ci[1], r.cj[1]);
        prim_t g01 = load_grad(m, G, axis, r.leaf[2], r.ci[2], r.cj[2]);
        prim_t g11 = load_grad(m, G, axis, r.leaf[3], r.ci[3], r.cj[3]);
        return {((g00.s + g10.s) * 0.5 + (g01.s + g11.s) * 0.5) * 0.5,
                ((g00.vx + g10.vx) * 0.5 + (g01.vx + g11.vx) * 0.5) * 0.5,
                ((g00.vy + g10.vy) * 0.5 + (g01.vy + g11.vy) * 0.5) * 0.5};
    }

    __device__ __forceinline__ prim_t grad_at(const mesh_dev_t& m, const double* G, int axis, int b, int i, int j)
    {
        return grad_from_ref(m, G, axis, resolve_cell(m, b, i, j));
    }

    /** P2 + P3 for the listed blocks: physical PLM gradients at the block's own spacing. */
    __global__ void __launch_bounds__(THREADS) general_gradients(
        mesh_dev_t mesh, const stage_t* __restrict__ stage_ptr, const int* __restrict__ list,
        const double* __restrict__ Uin, double* __restrict__ G)
    {
        const stage_t S = *stage_ptr;
        const int N = mesh.N, b = list[blockIdx.x];
        const double theta = S.theta, inv_h = 1.0 / mesh.spacing[b];
        const size_t base = size_t(mesh.gslot[b]) * N * N;

        for (int k = threadIdx.x; k < N * N; k += THREADS)
        {
            int i = k / N, j = k % N;
            prim_t c  = prim_at(mesh, Uin, b, i, j);
            prim_t xl = prim_at(mesh, Uin, b, i - 1, j), xr = prim_at(mesh, Uin, b, i + 1, j);
            prim_t yl = prim_at(mesh, Uin, b, i, j - 1), yr = prim_at(mesh, Uin, b, i, j + 1);
            G[0 * mesh.GS + base + k] = plm_diff(xl.s,  c.s,  xr.s,  theta) * inv_h;
            G[1 * mesh.GS + base + k] = plm_diff(xl.vx, c.vx, xr.vx, theta) * inv_h;
            G[2 * mesh.GS + base + k] = plm_diff(xl.vy, c.vy, xr.vy, theta) * inv_h;
            G[3 * mesh.GS + base + k] = plm_diff(yl.s,  c.s,  yr.s,  theta) * inv_h;
            G[4 * mesh.GS + base + k] = plm_diff(yl.vx, c.vx, yr.vx, theta) * inv_h;
            G[5 * mesh.GS + base + k] = plm_diff(yl.vy, c.vy, yr.vy, theta) * inv_h;
        }
    }

    /** The same in TX x TY tiles: the tile's primitives plus one guard layer go through shared memory once. */
    template<int TX, int TY>
    __global__ void __launch_bounds__(THREADS) general_gradients_tiled(
        mesh_dev_t mesh, const stage_t* __restrict__ stage_ptr, const int* __restrict__ list,
        const double* __restrict__ Uin, double* __restrict__ G)
    {
        __shared__ double P[3][TX + 2][TY + 2];
        const stage_t S = *stage_ptr;
        const int N = mesh.N;
        const int tiles_y = N / TY, tiles_per_block = (N / TX) * tiles_y;
        const int b = list[blockIdx.x / tiles_per_block], t = blockIdx.x % tiles_per_block;
        const int i0 = (t / tiles_y) * TX, j0 = (t % tiles_y) * TY;
        const double theta = S.theta, inv_h = 1.0 / mesh.spacing[b];
        const size_t base = size_t(mesh.gslot[b]) * N * N;

        for (int k = threadIdx.x; k < (TX + 2) * (TY + 2); k += THREADS)
        {
            const int li = k / (TY + 2), lj = k % (TY + 2);
            if ((li == 0 || li == TX + 1) && (lj == 0 || lj == TY + 1)) continue;      // corners are not part of the stencil
            const prim_t p = prim_at(mesh, Uin, b, i0 - 1 + li, j0 - 1 + lj);
            P[0][li][lj] = p.s; P[1][li][lj] = p.vx; P[2][li][lj] = p.vy;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < TX * TY; k += THREADS)
        {
            const int li = k / TY + 1, lj = k % TY + 1;
            const size_t c = base + size_t(i0 + li - 1) * N + (j0 + lj - 1);
            #pragma unroll
            for (int q = 0; q < 3; ++q)
            {
                G[q * mesh.GS + c]       = plm_diff(P[q][li - 1][lj], P[q][li][lj], P[q][li + 1][lj], theta) * inv_h;
                G[(3 + q) * mesh.GS + c] = plm_diff(P[q][li][lj - 1], P[q][li][lj], P[q][li][lj + 1], theta) * inv_h;
            }
        }
    }

    /**
     * The same for the two outermost cell layers along each side of the listed blocks only, one CTA per (block, side):
     * all that stage_strip<.., JUMP> reads from a neighbour (guard gradients are injected from / averaged over cells at
     * most two deep, and the fine faces of the flux correction touch the outermost layer).
     */
    __global__ void __launch_bounds__(128, 8) general_gradients_ring(
        mesh_dev_t mesh, const stage_t* __restrict__ stage_ptr, const int* __restrict__ list,
        const double* __restrict__ Uin, double* __restrict__ G)
    {
        // Region cell (r, a): r = 0..3 counts layers from the guard layer (r = 0) inwards, a = 0..N+1 runs along the side
        // from the guard cell before its first cell to the one after its last; primitives go through shared memory once.
        extern __shared__ double ring_P[];              // [3][4][N + 2]
        const stage_t S = *stage_ptr;
        const int N = mesh.N, W = N + 2, b = list[blockIdx.x >> 2], side = blockIdx.x & 3;
        const bool high = side & 1, along_x = side >= 2;        // along_x: the side runs along i (sides in y)
        const double theta = S.theta, inv_h = 1.0 / mesh.spacing[b];
        const size_t base = size_t(mesh.gslot[b]) * N * N;

        for (int k = threadIdx.x; k < 4 * W; k += 128)
        {
            const int r = k / W, a = k - r * W;
            const int n = high ? N - r : r - 1, t = a - 1;
            if (r == 0 && (t < 0 || t >= N)) continue;          // corners are not part of the stencil
            const prim_t p = along_x ? prim_at(mesh, Uin, b, t, n) : prim_at(mesh, Uin, b, n, t);
            ring_P[(0 * 4 + r) * W + a] = p.s; ring_P[(1 * 4 + r) * W + a] = p.vx; ring_P[(2 * 4 + r) * W + a] = p.vy;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < 2 * N; k += 128)
        {
            const int d = k / N, t = k - d * N, r = d + 1, a = t + 1;
            const int n = high ? N - r : r - 1;
            const size_t cell = base + (along_x ? size_t(t) * N + n : size_t(n) * N + t);
            #pragma unroll
            for (int q = 0; q < 3; ++q)
            {
                const double* Pq = ring_P + size_t(q) * 4 * W;
                const double c = Pq[r * W + a];
                const double below = Pq[(high ? r + 1 : r - 1) * W + a], above = Pq[(high ? r - 1 : r + 1) * W + a];
                const double gn = plm_diff(below, c, above, theta) * inv_h;                         // across the layers
                const double gt = plm_diff(Pq[r * W + a - 1], c, Pq[r * W + a + 1], theta) * inv_h; // along the side
                G[(along_x ? 3 + q : q) * mesh.GS + cell] = gn;
                G[(along_x ? q : 3 + q) * mesh.GS + cell] = gt;
            }
        }
    }

    /**
     * Flux (times face length) through face f (0..N) of block b along AXIS at transverse index k,
     * as block b computes it: block_fluxes_u (scheme.cpp:472-516).
     */
    template<int AXIS>
    __device__ void general_face_flux(const mesh_dev_t& m, const model_t& model, const stage_t& S,
        const double* U, const double* G, int b, int f, int k, double F[3])
    {
        const int N = m.N;
        const double* xv = m.xv + size_t(b) * (N + 1);
        const double* yv = m.yv + size_t(b) * (N + 1);
        int li = AXIS == 0 ? f - 1 : k, lj = AXIS == 0 ? k : f - 1;
        int ri = AXIS == 0 ? f : k,     rj = AXIS == 0 ? k : f;
        double x   = AXIS == 0 ? xv[f] : 0.5 * (xv[k] + xv[k + 1]);
        double y   = AXIS == 0 ? 0.5 * (yv[k] + yv[k + 1]) : yv[f];
        double len = AXIS == 0 ? yv[k + 1] - yv[k] : xv[k + 1] - xv[k];

        prim_t pl = prim_at(m, U, b, li, lj), pr = prim_at(m, U, b, ri, rj);
        prim_t gl = grad_at(m, G, AXIS, b, li, lj), gr = grad_at(m, G, AXIS, b, ri, rj);
        prim_t hl = grad_at(m, G, 1 - AXIS, b, li, lj), hr = grad_at(m, G, 1 - AXIS, b, ri, rj);
        eos_t e = eos_at_face(model, S, x, y);
        face_flux<AXIS>(e, pl, pr, gl, gr, hl.vx, hl.vy, hr.vx, hr.vy, 0.5 * m.spacing[b], 1.0, F);
        if (m.qmode) to_angmom_fluxes<AXIS>(model, x, y, F);
        F[0] *= len; F[1] *= len; F[2] *= len;
    }

    /** The same with correct_fluxes_{x,y} applied (scheme.cpp:614-720). */
    template<int AXIS>
    __device__ void general_face_flux_corrected(const mesh_dev_t& m, const model_t& model, const stage_t& S,
        const double* U, const double* G, int b, int f, int k, double F[3])
    {
        const int N = m.N;
        int side = f == 0 ? 2 * AXIS : (f == N ? 2 * AXIS + 1 : -1);

        if (side >= 0 && m.nbr[b * 4 + side].kind == 2)
        {
            // the neighbour region is refined: sum of the two fine faces, computed as the fine blocks do
            const face_nbr_dev_t nb = m.nbr[b * 4 + side];
            int near = side % 2 ? 0 : 1;                    // children adjacent to this face
            int fine_face = side % 2 ? 0 : N;
            double A[3], C[3];
            int k0 = 2 * k, k1 = 2 * k + 1;
            int c0 = AXIS == 0 ? nb.leaf[near + 2 * (k0 >= N)] : nb.leaf[(k0 >= N) + 2 * near];
            int c1 = AXIS == 0 ? nb.leaf[near + 2 * (k1 >= N)] : nb.leaf[(k1 >= N) + 2 * near];
            general_face_flux<AXIS>(m, model, S, U, G, c0, fine_face, k0 % N, A);
            general_face_flux<AXIS>(m, model, S, U, G, c1, fine_face, k1 % N, C);
            F[0] = A[0] + C[0]; F[1] = A[1] + C[1]; F[2] = A[2] + C[2];
            return;
        }
        general_face_flux<AXIS>(m, model, S, U, G, b, f, k, F);
    }

    /** P6-P8 + P11 for the listed blocks, one CTA per block. */
    __global__ void __launch_bounds__(THREADS) general_update(
        mesh_dev_t mesh, model_t model, const stage_t* __restrict__ stage_ptr, const int* __restrict__ list,
        const double* __restrict__ Uin, const double* __restrict__ G, const double* __restrict__ Un, double* __restrict__ Uout,
        double* __restrict__ partials, fail_dev_t* fail)
    {
        __shared__ double red[(THREADS / 32) * (NUM_SUMS + 1)];
        const stage_t S = *stage_ptr;
        const int N = mesh.N, b = list[blockIdx.x];
        const size_t FS = mesh.FS;
        const double* xv = mesh.xv + size_t(b) * (N + 1);
        const double* yv = mesh.yv + size_t(b) * (N + 1);
        const double h = mesh.spacing[b];

        double sums[NUM_SUMS];
        #pragma unroll
        for (int k = 0; k < NUM_SUMS; ++k) sums[k] = 0.0;
        double dtmin = 1e300;

        for (int k = threadIdx.x; k < N * N; k += THREADS)
        {
            int i = k / N, j = k % N;
            size_t c = size_t(b) * N * N + k;
            double Fl[3], Fr[3], Fb[3], Ft[3];
            general_face_flux_corrected<0>(mesh, model, S, Uin, G, b, i, j, Fl);
            general_face_flux_corrected<0>(mesh, model, S, Uin, G, b, i + 1, j, Fr);
            general_face_flux_corrected<1>(mesh, model, S, Uin, G, b, j, i, Fb);
            general_face_flux_corrected<1>(mesh, model, S, Uin, G, b, j + 1, i, Ft);

            double s = Uin[c], px = Uin[FS + c], py = Uin[2 * FS + c];
            double br = mesh.br[c];
            double u0s = mesh.U0[c], u0x = mesh.U0[FS + c], u0y = mesh.U0[2 * FS + c];
            double x = 0.5 * (xv[i] + xv[i + 1]), y = 0.5 * (yv[j] + yv[j + 1]);
            double dt_over_dA = S.dt / ((xv[i + 1] - xv[i]) * (yv[j + 1] - yv[j]));
            double src[3], y1, y2;
            if (mesh.qmode)
            {
                const prim_t p = load_prim(mesh, Uin, b, i, j);
                source_terms_q(model, S, x, y, s, px, py, p.vx, p.vy, u0s, u0x, u0y, br, src, sums, y1, y2);
            }
            else source_terms(model, S, x, y, s, px, py, u0s, u0x, u0y, br, src, sums, y1, y2);

            double n0 = s  - ((Fr[0] - Fl[0]) + (Ft[0] - Fb[0])) * dt_over_dA + src[0];
            double n1 = px - ((Fr[1] - Fl[1]) + (Ft[1] - Fb[1])) * dt_over_dA + src[1];
            double n2 = py - ((Fr[2] - Fl[2]) + (Ft[2] - Fb[2])) * dt_over_dA + src[2];

            if (n0 < 0.0) report_negative(fail, b, k, n0);

            if (S.combine)
            {
                double w = 1.0 - S.rk_b0;
                n0 = Un[c] * S.rk_b0 + n0 * w;
                n1 = Un[FS + c] * S.rk_b0 + n1 * w;
                n2 = Un[2 * FS + c] * S.rk_b0 + n2 * w;
            }
            Uout[c] = n0; Uout[FS + c] = n1; Uout[2 * FS + c] = n2;

            if (S.compute_dt)
            {
                double mx = n1, my = n2;
                if (mesh.qmode) angmom_to_linear(x, y, n1, n2, mx, my);
                dtmin = fmin(dtmin, h / max_wavespeed(model, S, x, y, y1, y2, n0, mx, my));
            }
        }
        reduce_and_store(red, sums, dtmin, h * h, partials + size_t(blockIdx.x) * ROW);
    }


    /**
     * P6-P8 + P11 for blocks at refinement jumps, one CTA per TX x TY tile: the same update as general_update, but the
     * tile's primitives (guard cells through prolongation / restriction, mesh_tree_operators.hpp:223-252) and the
     * gradients of general_gradients are staged in shared memory once, every face flux is computed once, and only the
     * faces on a block side whose neighbour is finer take the slow path (sum of the two fine fluxes, scheme.cpp:614-720).
     * Rows: one per tile, folded per block by finish_stage like the fused kernels' rows.
     */
    template<int TX, int TY>
    __global__ void __launch_bounds__(THREADS, 3) general_update_tiled(
        mesh_dev_t mesh, model_t model, const stage_t* __restrict__ stage_ptr, const int* __restrict__ list,
        const double* __restrict__ Uin, const double* __restrict__ G, const double* __restrict__ Un, double* __restrict__ Uout,
        double* __restrict__ partials, fail_dev_t* fail)
    {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        tile_t<TX, TY>& T = *reinterpret_cast<tile_t<TX, TY>*>(smem_raw);

        const stage_t S = *stage_ptr;
        const int N = mesh.N;
        const int tiles_y = N / TY, tiles_per_block = (N / TX) * tiles_y;
        const int b  = list[blockIdx.x / tiles_per_block];
        const int t  = blockIdx.x % tiles_per_block;
        const int i0 = (t / tiles_y) * TX, j0 = (t % tiles_y) * TY;
        const size_t FS = mesh.FS;
        const int tid = threadIdx.x;
        const double h = mesh.spacing[b];

        // which block sides border a finer neighbour (their faces take the flux-correction path): loaded now, used after the barrier
        const bool finer_lo_x = i0 == 0 && mesh.nbr[b * 4 + 0].kind == 2, finer_hi_x = i0 + TX == N && mesh.nbr[b * 4 + 1].kind == 2;
        const bool finer_lo_y = j0 == 0 && mesh.nbr[b * 4 + 2].kind == 2, finer_hi_y = j0 + TY == N && mesh.nbr[b * 4 + 3].kind == 2;

        // ---- tile + 1 guard layer (no corners: a face only needs its two cells): primitives and physical gradients
        for (int k = tid; k < (TX + 2) * (TY + 2); k += THREADS)
        {
            const int li = k / (TY + 2), lj = k % (TY + 2);         // region cell <-> block cell (i0 - 1 + li, j0 - 1 + lj)
            const bool edge_i = li == 0 || li == TX + 1, edge_j = lj == 0 || lj == TY + 1;
            if (edge_i && edge_j) continue;
            const int gi = i0 - 1 + li, gj = j0 - 1 + lj;
            const cell_ref_t ref = resolve_cell(mesh, b, gi, gj);       // once for the primitive and both gradients
            const prim_t p = prim_from_ref(mesh, Uin, ref);
            const prim_t gx = grad_from_ref(mesh, G, 0, ref), gy = grad_from_ref(mesh, G, 1, ref);
            T.P[0][li][lj] = p.s;  T.P[1][li][lj] = p.vx;  T.P[2][li][lj] = p.vy;
            T.G[0][li][lj] = gx.s; T.G[1][li][lj] = gx.vx; T.G[2][li][lj] = gx.vy;
            T.G[3][li][lj] = gy.s; T.G[4][li][lj] = gy.vx; T.G[5][li][lj] = gy.vy;
        }
        if (tid <= TX) T.xv[tid] = mesh.xv[size_t(b) * (N + 1) + i0 + tid];
        if (tid >= 64 && tid - 64 <= TY) T.yv[tid - 64] = mesh.yv[size_t(b) * (N + 1) + j0 + tid - 64];
        __syncthreads();

        // ---- fluxes (times face length, as block_fluxes_u, scheme.cpp:472-516)

        auto x_face = [&] (int li, int lj)      // between tile cells (li - 1, lj) and (li, lj), 0 <= li <= TX
        {
            double F[3];
            if ((li == 0 && finer_lo_x) || (li == TX && finer_hi_x)) general_face_flux_corrected<0>(mesh, model, S, Uin, G, b, i0 + li, j0 + lj, F);
            else
            {
                const eos_t e = eos_at_face(model, S, T.xv[li], 0.5 * (T.yv[lj] + T.yv[lj + 1]));
                const prim_t pl = {T.P[0][li][lj + 1], T.P[1][li][lj + 1], T.P[2][li][lj + 1]}, pr = {T.P[0][li + 1][lj + 1], T.P[1][li + 1][lj + 1], T.P[2][li + 1][lj + 1]};
                const prim_t gl = {T.G[0][li][lj + 1], T.G[1][li][lj + 1], T.G[2][li][lj + 1]}, gr = {T.G[0][li + 1][lj + 1], T.G[1][li + 1][lj + 1], T.G[2][li + 1][lj + 1]};
                face_flux<0>(e, pl, pr, gl, gr, T.G[4][li][lj + 1], T.G[5][li][lj + 1], T.G[4][li + 1][lj + 1], T.G[5][li + 1][lj + 1], 0.5 * h, 1.0, F);
                if (mesh.qmode) to_angmom_fluxes<0>(model, T.xv[li], 0.5 * (T.yv[lj] + T.yv[lj + 1]), F);
                const double len = T.yv[lj + 1] - T.yv[lj];
                F[0] *= len; F[1] *= len; F[2] *= len;
            }
            T.Fx[0][li][lj] = F[0]; T.Fx[1][li][lj] = F[1]; T.Fx[2][li][lj] = F[2];
        };
        auto y_face = [&] (int li, int lj)      // between tile cells (li, lj - 1) and (li, lj), 0 <= lj <= TY
        {
            double F[3];
            if ((lj == 0 && finer_lo_y) || (lj == TY && finer_hi_y)) general_face_flux_corrected<1>(mesh, model, S, Uin, G, b, j0 + lj, i0 + li, F);
            else
            {
                const eos_t e = eos_at_face(model, S, 0.5 * (T.xv[li] + T.xv[li + 1]), T.yv[lj]);
                const prim_t pl = {T.P[0][li + 1][lj], T.P[1][li + 1][lj], T.P[2][li + 1][lj]}, pr = {T.P[0][li + 1][lj + 1], T.P[1][li + 1][lj + 1], T.P[2][li + 1][lj + 1]};
                const prim_t gl = {T.G[3][li + 1][lj], T.G[4][li + 1][lj], T.G[5][li + 1][lj]}, gr = {T.G[3][li + 1][lj + 1], T.G[4][li + 1][lj + 1], T.G[5][li + 1][lj + 1]};
                face_flux<1>(e, pl, pr, gl, gr, T.G[1][li + 1][lj], T.G[2][li + 1][lj], T.G[1][li + 1][lj + 1], T.G[2][li + 1][lj + 1], 0.5 * h, 1.0, F);
                if (mesh.qmode) to_angmom_fluxes<1>(model, 0.5 * (T.xv[li] + T.xv[li + 1]), T.yv[lj], F);
                const double len = T.xv[li + 1] - T.xv[li];
                F[0] *= len; F[1] *= len; F[2] *= len;
            }
            T.Fy[0][li][lj] = F[0]; T.Fy[1][li][lj] = F[1]; T.Fy[2][li][lj] = F[2];
        };
        for (int k = tid; k < TX * TY; k += THREADS)
        {
            x_face(k / TY, k % TY);
            y_face(k / TY, k % TY);
        }
        for (int k = tid; k < TX + TY; k += THREADS)
        {
            if (k < TY) x_face(TX, k); else y_face(k - TY, TY);
        }
        __syncthreads();

        // ---- update (block_update_u, scheme.cpp:568-587), validation, CFL estimate
        double sums[NUM_SUMS];
        #pragma unroll
        for (int k = 0; k < NUM_SUMS; ++k) sums[k] = 0.0;
        double dtmin = 1e300;

        for (int k = tid; k < TX * TY; k += THREADS)
        {
            const int li = k / TY, lj = k % TY;
            const size_t c = (size_t(b) * N + (i0 + li)) * N + (j0 + lj);
            const double s = Uin[c], px = Uin[FS + c], py = Uin[2 * FS + c];
            const double br = mesh.br[c];
            const double u0s = mesh.U0[c], u0x = mesh.U0[FS + c], u0y = mesh.U0[2 * FS + c];
            const double x = 0.5 * (T.xv[li] + T.xv[li + 1]), y = 0.5 * (T.yv[lj] + T.yv[lj + 1]);
            const double dt_over_dA = S.dt / ((T.xv[li + 1] - T.xv[li]) * (T.yv[lj + 1] - T.yv[lj]));
            double src[3], y1, y2;
            if (mesh.qmode) source_terms_q(model, S, x, y, s, px, py, T.P[1][li + 1][lj + 1], T.P[2][li + 1][lj + 1], u0s, u0x, u0y, br, src, sums, y1, y2);
            else source_terms(model, S, x, y, s, px, py, u0s, u0x, u0y, br, src, sums, y1, y2);

            double n0 = s  - ((T.Fx[0][li + 1][lj] - T.Fx[0][li][lj]) + (T.Fy[0][li][lj + 1] - T.Fy[0][li][lj])) * dt_over_dA + src[0];
            double n1 = px - ((T.Fx[1][li + 1][lj] - T.Fx[1][li][lj]) + (T.Fy[1][li][lj + 1] - T.Fy[1][li][lj])) * dt_over_dA + src[1];
            double n2 = py - ((T.Fx[2][li + 1][lj] - T.Fx[2][li][lj]) + (T.Fy[2][li][lj + 1] - T.Fy[2][li][lj])) * dt_over_dA + src[2];

            if (n0 < 0.0) report_negative(fail, b, (i0 + li) * N + j0 + lj, n0);

            if (S.combine)
            {
                const double w = 1.0 - S.rk_b0;
                n0 = Un[c] * S.rk_b0 + n0 * w;
                n1 = Un[FS + c] * S.rk_b0 + n1 * w;
                n2 = Un[2 * FS + c] * S.rk_b0 + n2 * w;
            }
            Uout[c] = n0; Uout[FS + c] = n1; Uout[2 * FS + c] = n2;

            if (S.compute_dt) dtmin = fmin(dtmin, h / max_wavespeed(model, S, x, y, y1, y2, n0, n1, n2));
        }
        reduce_and_store(T.red, sums, dtmin, h * h, partials + size_t(blockIdx.x) * ROW);
    }


}
#include "stage_strip.cuh"      // after the any-tree helpers: its JUMP variant fetches guard cells through them
namespace
{
    // =======================================================================
    // Stand-alone CFL pass, row folding, layout changes
    // =======================================================================

    /** maximum_timestep (scheme.cpp:1107-1126): per-CTA min of spacing / max wavespeed. */
    __global__ void __launch_bounds__(THREADS) max_timestep_kernel(
        mesh_dev_t mesh, model_t model, const stage_t* __restrict__ stage_ptr, const double* __restrict__ U, double* __restrict__ partials)
    {
        __shared__ double red[THREADS / 32];
        const stage_t S = *stage_ptr;
        const int N = mesh.N, b = blockIdx.x;
        const double* xv = mesh.xv + size_t(b) * (N + 1);
        const double* yv = mesh.yv + size_t(b) * (N + 1);
        const double h = mesh.spacing[b];
        double dtmin = 1e300;

        for (int k = threadIdx.x; k < N * N; k += THREADS)
        {
            int i = k / N, j = k % N;
            size_t c = size_t(b) * N * N + k;
            double x = 0.5 * (xv[i] + xv[i + 1]), y = 0.5 * (yv[j] + yv[j + 1]);
            double y1, y2;
            sound_speed_squared(model, S, x, y, y1, y2);
            double mx = U[mesh.FS + c], my = U[2 * mesh.FS + c];
            if (mesh.qmode) angmom_to_linear(x, y, mx, my, mx, my);      // (the reference itself needs fixed_dt = 1 with conserved_q)
            dtmin = fmin(dtmin, h / max_wavespeed(model, S, x, y, y1, y2, U[c], mx, my));
        }
        dtmin = warp_min(dtmin);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dtmin;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            for (int w = 1; w < THREADS / 32; ++w) dtmin = fmin(dtmin, red[w]);
            double* row = partials + size_t(b) * ROW;
            for (int k = 0; k < NUM_SUMS; ++k) row[k] = 0.0;
            row[NUM_SUMS] = dtmin;
        }
    }

    __device__ void fill_stage(stage_t& st, double time, double dt, double theta, const two_body_t& b, double rk_b0, int combine, int compute_dt)
    {
        st.time = time; st.dt = dt; st.theta = theta;
        st.x1 = b.body1.x; st.y1 = b.body1.y; st.m1 = b.body1.mass; st.vx1 = b.body1.vx; st.vy1 = b.body1.vy;
        st.x2 = b.body2.x; st.y2 = b.body2.y; st.m2 = b.body2.mass; st.vx2 = b.body2.vx; st.vy2 = b.body2.vy;
        st.rk_b0 = rk_b0; st.combine = combine; st.compute_dt = compute_dt;
    }

    /** Everything of a stage input except dt (written later, when the CFL reduction is known). */
    __device__ void fill_stage_but_dt(stage_t& st, double time, double theta, const two_body_t& b, double rk_b0, int combine, int compute_dt)
    {
        st.time = time; st.theta = theta;
        st.x1 = b.body1.x; st.y1 = b.body1.y; st.m1 = b.body1.mass; st.vx1 = b.body1.vx; st.vy1 = b.body1.vy;
        st.x2 = b.body2.x; st.y2 = b.body2.y; st.m2 = b.body2.mass; st.vx2 = b.body2.vx; st.vy2 = b.body2.vy;
        st.rk_b0 = rk_b0; st.combine = combine; st.compute_dt = compute_dt;
    }

    /**
     * The slow half of prepare_next, off the critical path (side stream): body positions (a Kepler solve and a
     * handful of divisions, ~9 us for one thread) for stages whose TIME is already known.  `src` is the first
     * stage of a step with (time, dt) = (t, dt): the step's second stage `second` runs at t + dt and the first
     * stage of the step after it, `following`, at t/2 + ((t + dt) + dt)/2 (scheme.cpp:1036, 1055).  Either may be null.
     */
    __global__ void prepare_positions(step_config_t cfg, const stage_t* __restrict__ src, stage_t* second, stage_t* following)
    {
        const double t = src->time, dt = src->dt;
        if (threadIdx.x == 0 && second)
        {
            const double tb = t + dt;
            fill_stage_but_dt(*second, tb, cfg.theta, two_body_state(cfg.elements, tb), 0.5, 1, ! cfg.fixed_dt);
        }
        if (threadIdx.x == 1 && following)
        {
            const double tn = t * 0.5 + ((t + dt) + dt) * 0.5;
            fill_stage_but_dt(*following, tn, cfg.theta, two_body_state(cfg.elements, tn), 0.0, 0, 0);
        }
    }

    // =======================================================================
    // Peer-memory guard-zone exchange (NVLink loads / stores, no NCCL in the step loop)
    // =======================================================================
    /**
     * prepare_next for several ranks without NCCL (called by the >= 128 threads of one CTA): deliver this rank's
     * two stage results to every rank's mailbox, wait for everybody else's, fold them in rank order (every rank
     * gets the same bits) and write time and dt of the next step's stages.
     */
    __device__ void peer_prepare(const stage_result_t* __restrict__ local, const peer_table_t& peers, int me, int nranks,
        int slot_stride, int slot_a, int slot_b, unsigned long long counter,
        const step_config_t& cfg, const stage_t* __restrict__ current_a, stage_t* next_a, stage_t* next_b, stage_result_t* host_results,
        unsigned long long* clock_words = nullptr)
    {
        __shared__ double dt_min_b;
        constexpr int words = sizeof(stage_result_t) / sizeof(double);

        for (int k = threadIdx.x; k < nranks * 2 * words; k += blockDim.x)
        {
            const int p = k / (2 * words), slot = (k / words) % 2 ? slot_b : slot_a, w = k % words;
            reinterpret_cast<double*>(peers.results[p] + size_t(me) * slot_stride + slot)[w] = __ldcg(reinterpret_cast<const double*>(local + slot) + w);
        }
        __threadfence_system();
        __syncthreads();
        unsigned long long t_wait = 0;
        if (threadIdx.x == 0 && clock_words) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_wait));
        if (threadIdx.x < nranks)
        {
            if (threadIdx.x != me) store_release_sys(peers.result_flag[threadIdx.x] + me, counter);
            if (threadIdx.x != me) bounded_wait_sys(peers.result_flag[me] + threadIdx.x, counter, peers, me, threadIdx.x);
        }
        __syncthreads();
        if (threadIdx.x == 0 && clock_words)
        {
            // how long this rank waited for the slowest rank's results (bench.py: exchange.result_exchange_us)
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            clock_words[0] += t1 - t_wait;
            clock_words[1] += 1;
        }

        const int k = threadIdx.x;
        if (k < 2)
        {
            const int slot = k == 0 ? slot_a : slot_b;
            stage_result_t r = stage_result_t();
            r.dt_min = 1e300;
            for (int p = 0; p < nranks; ++p)
            {
                const double* q = reinterpret_cast<const double*>(peers.results[me] + size_t(p) * slot_stride + slot);
                for (int c = 0; c < 16; ++c) r.sums[c] += __ldcg(q + c);
                r.work[0] += __ldcg(q + 16);
                r.work[1] += __ldcg(q + 17);
                r.dt_min = dmin(r.dt_min, __ldcg(q + 18));
                r.num_negative += __ldcg(reinterpret_cast<const unsigned int*>(q + 19));
            }
            r.pad = static_cast<unsigned int>(load_acquire_sys(peers.abort_word[me]));     // non-zero: some rank gave up waiting (bounded_wait_sys)
            host_results[slot] = r;
            if (k == 1) dt_min_b = r.dt_min;
        }
        __syncthreads();
        if (k < 2)
        {
            const double t = current_a->time, dt = current_a->dt;
            const double t_next = t * 0.5 + ((t + dt) + dt) * 0.5;
            const double dt_next = cfg.fixed_dt ? cfg.recommended_time_step : cfg.cfl_number * dt_min_b;
            if (k == 0) { next_a->time = t_next; next_a->dt = dt_next; }        // positions: prepare_positions, off the critical path
            else        { next_b->time = t_next + dt_next; next_b->dt = dt_next; }
        }
    }

    /**
     * Fold the stage kernels' rows in a fixed order (deterministic) and publish the stage result.
     * Rows [0, num_fused) are regular blocks: their `tpb` tile rows (written by stage_strip / stage_fused,
     * no fences or tickets in those kernels) are first folded, in tile order, into one row per block;
     * rows [num_fused, num_rows) are the any-tree path's blocks, `gtpb` tile rows each (1: one row per block).
     * The reference evaluates the work done on each body PER BLOCK from that block's accreted
     * mass and momentum -- a non-linear function -- and then sums over blocks
     * (scheme.cpp:407-408, 829-830), so the same is done here from the per-block rows.
     * CTA c handles FINISH_ROWS_PER_CTA block rows; the last CTA to finish folds the CTA rows in CTA order.
     */
    __global__ void __launch_bounds__(FINISH_THREADS) finish_stage(const double* tile_rows, int num_fused, int tpb,
        const double* general_rows, int gtpb, int num_rows, double* cta_rows, int* ticket, const stage_t* __restrict__ stage_ptr,
        fail_dev_t* fail, stage_result_t* result, prepare_args_t prep)
    {
        __shared__ double dt_min_all;
        __shared__ double srow[FINISH_ROWS_PER_CTA][ROW];
        __shared__ double wred[2][FINISH_ROWS_PER_CTA];
        __shared__ double fin[FINISH_THREADS / 32][32];
        __shared__ int is_last;
        const stage_t S = *stage_ptr;
        const int r0 = blockIdx.x * FINISH_ROWS_PER_CTA, n = min(FINISH_ROWS_PER_CTA, num_rows - r0);

        // (A) one row per block
        for (int idx = threadIdx.x; idx < n * ROW; idx += FINISH_THREADS)
        {
            const int r = idx / ROW, k = idx % ROW, R = r0 + r;
            if (k > NUM_SUMS) continue;
            double v;
            if (R < num_fused)
            {
                const double* rows = tile_rows + size_t(R) * tpb * ROW + k;
                v = k == NUM_SUMS ? 1e300 : 0.0;
                for (int t = 0; t < tpb; ++t)
                {
                    double p = __ldcg(rows + size_t(t) * ROW);
                    v = k == NUM_SUMS ? dmin(v, p) : v + p;
                }
            }
            else
            {
                // blocks of the any-tree path: gtpb tile rows each (general_update_tiled), or one row (general_update)
                const double* rows = general_rows + size_t(R - num_fused) * gtpb * ROW + k;
                v = __ldcg(rows);
                for (int t = 1; t < gtpb; ++t)
                {
                    double p = __ldcg(rows + size_t(t) * ROW);
                    v = k == NUM_SUMS ? dmin(v, p) : v + p;
                }
            }
            srow[r][k] = v;
        }
        __syncthreads();

        // (B) fold the CTA's rows in row order; block-wise work integrals
        double* mine = cta_rows + size_t(blockIdx.x) * ROW;
        if (threadIdx.x <= NUM_SUMS)
        {
            const int k = threadIdx.x;
            double v = k == NUM_SUMS ? 1e300 : 0.0;
            for (int r = 0; r < n; ++r) v = k == NUM_SUMS ? dmin(v, srow[r][k]) : v + srow[r][k];
            mine[k] = v;
        }
        else if (threadIdx.x >= 32 && threadIdx.x < 32 + 2 * FINISH_ROWS_PER_CTA)
        {
            const int r = (threadIdx.x - 32) >> 1, k = threadIdx.x & 1;
            double w = 0.0;
            if (r < n)
            {
                const double dm = srow[r][ACC_MASS + k], dpx = srow[r][ACC_PX + k], dpy = srow[r][ACC_PY + k];
                if (dm != 0.0 || dpx != 0.0 || dpy != 0.0)
                {
                    const double M0 = k ? S.m2 : S.m1, px0 = (k ? S.vx2 : S.vx1) * M0, py0 = (k ? S.vy2 : S.vy1) * M0;
                    const double M1 = M0 + dm * S.dt, px1 = px0 + dpx * S.dt, py1 = py0 + dpy * S.dt;
                    w = ((px1 * px1 + py1 * py1) / M1 - (px0 * px0 + py0 * py0) / M0) * 0.5;
                }
            }
            wred[k][r] = w;
        }
        __syncthreads();
        if (threadIdx.x < 2)
        {
            double w = 0.0;
            for (int r = 0; r < n; ++r) w += wred[threadIdx.x][r];
            mine[NUM_SUMS + 1 + threadIdx.x] = w;
        }
        if (threadIdx.x < 32) __threadfence();      // the writers of `mine` all sit in warp 0
        __syncthreads();
        if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1) == int(gridDim.x) - 1;
        __syncthreads();
        if (! is_last) return;
        __threadfence();

        // (C) CTA rows: group g folds rows g, g + 8, ... in order, then the eight groups are folded in order
        const int col = threadIdx.x % 32, grp = threadIdx.x / 32, ngrp = FINISH_THREADS / 32;
        const bool is_min = col == NUM_SUMS;
        double v = is_min ? 1e300 : 0.0;
        if (col < ROW - 1)
        {
            for (int c = grp; c < int(gridDim.x); c += ngrp)
            {
                double p = __ldcg(cta_rows + size_t(c) * ROW + col);
                v = is_min ? dmin(v, p) : v + p;
            }
        }
        fin[grp][col] = v;
        __syncthreads();
        if (threadIdx.x < ROW - 1)
        {
            const int k = threadIdx.x;
            double f = fin[0][k];
            for (int g = 1; g < ngrp; ++g) f = k == NUM_SUMS ? dmin(f, fin[g][k]) : f + fin[g][k];
            if (k < NUM_SUMS) result->sums[k] = f;
            else if (k == NUM_SUMS) { result->dt_min = f; dt_min_all = f; }
            else result->work[k - NUM_SUMS - 1] = f;
        }
        if (threadIdx.x == 64)
        {
            result->num_negative = fail->count;
            fail->pad = fail->count;    // how many entries of the list belong to this launch
            fail->count = 0;            // ready for the next launch that uses this slot
            *ticket = 0;
        }
        if (! prep.enabled) return;
        if (prep.enabled == 2)
        {
            __threadfence();            // this launch's own result (written above) is read back through global memory
            __syncthreads();
            peer_prepare(prep.local, prep.peers, prep.me, prep.nranks, prep.slot_stride, prep.slot_a, prep.slot_b, prep.counter,
                         prep.cfg, prep.current_a, prep.next_a, prep.next_b, prep.host_results, prep.clock_words);
            return;
        }

        // stage inputs of the next step (see prepare_next): one thread per stage
        __syncthreads();
        if (threadIdx.x == 96 || threadIdx.x == 97)
        {
            const double t = prep.current_a->time, dt = prep.current_a->dt;
            const double t_next = t * 0.5 + ((t + dt) + dt) * 0.5;
            const double dt_next = prep.cfg.fixed_dt ? prep.cfg.recommended_time_step : prep.cfg.cfl_number * dt_min_all;
            // time and dt only: the body positions of next_a were prepared a step ago, those of next_b follow on the
            // side stream while next_a runs (prepare_positions)
            if (threadIdx.x == 96) { prep.next_a->time = t_next; prep.next_a->dt = dt_next; }
            else { prep.next_b->time = t_next + dt_next; prep.next_b->dt = dt_next; }
        }
    }

    /**
     * finish_stage for up to FINISH_CLUSTER_MAX_ROWS blocks as ONE CLUSTER of eight CTAs (thread-block cluster, distributed
     * shared memory): every CTA folds its share of the blocks -- tile rows -> block row (tile order), block-wise work integral,
     * column sums by 32 interleaved groups -- then a hardware cluster barrier replaces the __threadfence + ticket of the
     * multi-CTA version, and CTA 0 folds the eight partial rows through DSMEM in rank order, publishes the result and writes
     * the next step's time and dt.  This kernel sits on the critical path of every step (stage b -> finish -> next stage a);
     * eight SMs give it the memory-level parallelism one CTA lacks (the tile rows come from L2, ~2000 cycles away).
     */
    constexpr int FINISH_CLUSTER = 8, FINISH_CLUSTER_THREADS = 1024, FINISH_CLUSTER_ROWS = 128;
    constexpr int FINISH_CLUSTER_MAX_ROWS = FINISH_CLUSTER * FINISH_CLUSTER_ROWS;

    __global__ void __cluster_dims__(FINISH_CLUSTER, 1, 1) __launch_bounds__(FINISH_CLUSTER_THREADS) finish_stage_cluster(
        const double* tile_rows, int num_fused, int tpb, const double* general_rows, int gtpb, int num_rows,
        const stage_t* __restrict__ stage_ptr, fail_dev_t* fail, stage_result_t* result, prepare_args_t prep)
    {
        namespace cg = cooperative_groups;
        __shared__ double srows[FINISH_CLUSTER_ROWS][ROW];
        __shared__ double part[32][ROW];
        __shared__ double mine[ROW];
        __shared__ double dt_min_all;
        auto cluster = cg::this_cluster();
        const int tid = threadIdx.x, rank = int(cluster.block_rank());
        const int per = (num_rows + FINISH_CLUSTER - 1) / FINISH_CLUSTER;
        const int r0 = min(num_rows, rank * per), n = min(num_rows, r0 + per) - r0;

        // (A) one row per block, tiles in tile order, eight loads in flight at a time
        for (int idx = tid; idx < n * ROW; idx += FINISH_CLUSTER_THREADS)
        {
            const int r = idx / ROW, k = idx % ROW, R = r0 + r;
            if (k > NUM_SUMS) continue;
            const bool fused = R < num_fused;
            const int nt = fused ? tpb : gtpb;
            const double* rows = (fused ? tile_rows + size_t(R) * tpb * ROW : general_rows + size_t(R - num_fused) * gtpb * ROW) + k;
            double v = k == NUM_SUMS ? 1e300 : 0.0;
            for (int t0 = 0; t0 < nt; t0 += 8)
            {
                double p[8];
                #pragma unroll
                for (int t = 0; t < 8; ++t) p[t] = t0 + t < nt ? __ldcg(rows + size_t(t0 + t) * ROW) : (k == NUM_SUMS ? 1e300 : 0.0);
                #pragma unroll
                for (int t = 0; t < 8; ++t) v = k == NUM_SUMS ? dmin(v, p[t]) : (t0 + t < nt ? v + p[t] : v);
            }
            srows[r][k] = v;
        }
        __syncthreads();

        // block-wise work integral (scheme.cpp:363-374, 407-408) into the two spare columns of the block's row
        const stage_t S = *stage_ptr;
        for (int idx = tid; idx < 2 * n; idx += FINISH_CLUSTER_THREADS)
        {
            const int r = idx >> 1, k = idx & 1;
            const double dm = srows[r][ACC_MASS + k], dpx = srows[r][ACC_PX + k], dpy = srows[r][ACC_PY + k];
            double w = 0.0;
            if (dm != 0.0 || dpx != 0.0 || dpy != 0.0)
            {
                const double M0 = k ? S.m2 : S.m1, px0 = (k ? S.vx2 : S.vx1) * M0, py0 = (k ? S.vy2 : S.vy1) * M0;
                const double M1 = M0 + dm * S.dt, px1 = px0 + dpx * S.dt, py1 = py0 + dpy * S.dt;
                w = ((px1 * px1 + py1 * py1) / M1 - (px0 * px0 + py0 * py0) / M0) * 0.5;
            }
            srows[r][NUM_SUMS + 1 + k] = w;
        }
        __syncthreads();

        // (B) this CTA's columns: group g folds rows g, g + 32, ... in order, then the 32 groups are folded in order
        {
            const int col = tid % 32, grp = tid / 32;
            if (col < ROW - 1)
            {
                const bool is_min = col == NUM_SUMS;
                double v = is_min ? 1e300 : 0.0;
                for (int r = grp; r < n; r += 32)
                {
                    double p = srows[r][col];
                    v = is_min ? dmin(v, p) : v + p;
                }
                part[grp][col] = v;
            }
        }
        __syncthreads();
        if (tid < ROW - 1)
        {
            const int k = tid;
            double f = part[0][k];
            for (int g = 1; g < 32; ++g) f = k == NUM_SUMS ? dmin(f, part[g][k]) : f + part[g][k];
            mine[k] = f;
        }
        cluster.sync();

        // (C) CTA 0: the eight partial rows through distributed shared memory, in rank order
        if (rank == 0)
        {
            if (tid < ROW - 1)
            {
                const int k = tid;
                double f = mine[k];
                for (int c = 1; c < FINISH_CLUSTER; ++c)
                {
                    const double p = cluster.map_shared_rank(mine, c)[k];
                    f = k == NUM_SUMS ? dmin(f, p) : f + p;
                }
                if (k < NUM_SUMS) result->sums[k] = f;
                else if (k == NUM_SUMS) { result->dt_min = f; dt_min_all = f; }
                else result->work[k - NUM_SUMS - 1] = f;
            }
            if (tid == 64)
            {
                result->num_negative = fail->count;
                fail->pad = fail->count;
                fail->count = 0;
            }
        }
        cluster.sync();         // the other CTAs' shared memory stays alive until CTA 0 has read it
        if (rank != 0 || ! prep.enabled) return;
        if (prep.enabled == 2)
        {
            __threadfence();
            __syncthreads();
            peer_prepare(prep.local, prep.peers, prep.me, prep.nranks, prep.slot_stride, prep.slot_a, prep.slot_b, prep.counter,
                         prep.cfg, prep.current_a, prep.next_a, prep.next_b, prep.host_results, prep.clock_words);
            return;
        }
        if (tid == 96 || tid == 97)
        {
            const double t = prep.current_a->time, dt = prep.current_a->dt;
            const double t_next = t * 0.5 + ((t + dt) + dt) * 0.5;
            const double dt_next = prep.cfg.fixed_dt ? prep.cfg.recommended_time_step : prep.cfg.cfl_number * dt_min_all;
            if (tid == 96) { prep.next_a->time = t_next; prep.next_a->dt = dt_next; }
            else { prep.next_b->time = t_next + dt_next; prep.next_b->dt = dt_next; }
        }
    }

    /**
     * End of an RK2 step, on the device: fold the two stage results over the ranks (rank order, so every
     * rank gets the same bits), publish them to the host, and write the stage inputs of the NEXT step --
     * dt = cfl * min(spacing / wavespeed) (subprog_binary.cpp:281-283), time = t/2 + ((t + dt) + dt)/2
     * (scheme.cpp:1036, 1055), body positions from compute_two_body_state (scheme.cpp:814) -- so that
     * the host can queue the next step without waiting for this one.
     */
    __global__ void prepare_next(const stage_result_t* __restrict__ gathered, int nranks, int slot_stride, int slot_a, int slot_b,
        step_config_t cfg, const stage_t* __restrict__ current, stage_t* next_a, stage_t* next_b, stage_result_t* host_results)
    {
        // two threads: one per stage result / per stage of the next step
        __shared__ double dt_min_b;
        const int k = threadIdx.x;
        if (k >= 2) return;
        const int slot = k == 0 ? slot_a : slot_b;
        stage_result_t r = stage_result_t();
        r.dt_min = 1e300;

        for (int p = 0; p < nranks; ++p)
        {
            const stage_result_t& q = gathered[size_t(p) * slot_stride + slot];
            for (int c = 0; c < 16; ++c) r.sums[c] += q.sums[c];
            r.work[0] += q.work[0];
            r.work[1] += q.work[1];
            r.dt_min = dmin(r.dt_min, q.dt_min);
            r.num_negative += q.num_negative;
        }
        host_results[slot] = r;
        if (k == 1) dt_min_b = r.dt_min;
        __syncwarp(0x3);

        const double t = current[slot_a].time, dt = current[slot_a].dt;
        const double t_next = t * 0.5 + ((t + dt) + dt) * 0.5;
        const double dt_next = cfg.fixed_dt ? cfg.recommended_time_step : cfg.cfl_number * dt_min_b;
        if (k == 0) fill_stage(*next_a, t_next, dt_next, cfg.theta, two_body_state(cfg.elements, t_next), 0.0, 0, 0);
        else        fill_stage(*next_b, t_next + dt_next, dt_next, cfg.theta, two_body_state(cfg.elements, t_next + dt_next), 0.5, 1, ! cfg.fixed_dt);
    }

    /** One strip / corner of a block in the guard-zone exchange between ranks (partition.hpp). */
    /**
     * disk_mass and disk_angular_momentum of the time series (subprog_binary_diagnostics.cpp:19-41): per block the
     * sums of sigma dA and (x py - y px) dA, folded in a fixed order; the host adds the blocks in tree order.
     */
    __global__ void __launch_bounds__(THREADS) disk_totals_kernel(mesh_dev_t mesh, const double* __restrict__ U, double* __restrict__ out)
    {
        __shared__ double red[2][THREADS / 32];
        const int N = mesh.N, b = blockIdx.x;
        const double* xv = mesh.xv + size_t(b) * (N + 1);
        const double* yv = mesh.yv + size_t(b) * (N + 1);
        double m = 0.0, l = 0.0;

        for (int k = threadIdx.x; k < N * N; k += THREADS)
        {
            int i = k / N, j = k % N;
            size_t c = size_t(b) * N * N + k;
            double x = 0.5 * (xv[i] + xv[i + 1]), y = 0.5 * (yv[j] + yv[j + 1]);
            double dA = (xv[i + 1] - xv[i]) * (yv[j + 1] - yv[j]);
            m += U[c] * dA;
            l += (mesh.qmode ? U[2 * mesh.FS + c] : x * U[2 * mesh.FS + c] - y * U[mesh.FS + c]) * dA;     // Lz is the third component of conserved_q
        }
        m = warp_sum(m); l = warp_sum(l);
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = m; red[1][threadIdx.x >> 5] = l; }
        __syncthreads();
        if (threadIdx.x < 2)
        {
            double v = 0.0;
            for (int w = 0; w < THREADS / 32; ++w) v += red[threadIdx.x][w];
            out[2 * b + threadIdx.x] = v;
        }
    }

    /**
     * diagnostic_fields (subprog_binary_diagnostics.cpp:48-82): sigma, v_r = v . rhat, v_phi = v . phihat per cell,
     * block major [B][3][NN] for the writer.
     */
    __global__ void diagnostic_fields_kernel(mesh_dev_t mesh, const double* __restrict__ U, double* __restrict__ out, int BO)
    {
        const int N = mesh.N;
        const size_t NN = size_t(N) * N, n = size_t(BO) * NN;
        for (size_t k = blockIdx.x * size_t(blockDim.x) + threadIdx.x; k < n; k += size_t(gridDim.x) * blockDim.x)
        {
            const size_t b = k / NN, cell = k % NN;
            const int i = int(cell / N), j = int(cell % N);
            const double* xv = mesh.xv + b * (N + 1);
            const double* yv = mesh.yv + b * (N + 1);
            const double xc = (xv[i] + xv[i + 1]) * 0.5, yc = (yv[j] + yv[j + 1]) * 0.5;
            const double rc = sqrt(xc * xc + yc * yc);
            const double sigma = U[k];
            double vx = U[mesh.FS + k] / sigma, vy = U[2 * mesh.FS + k] / sigma;
            if (mesh.qmode) angmom_to_linear(xc, yc, vx, vy, vx, vy);
            out[(b * 3 + 0) * NN + cell] = sigma;
            out[(b * 3 + 1) * NN + cell] = vx * (xc / rc) + vy * (yc / rc);
            out[(b * 3 + 2) * NN + cell] = vx * (-yc / rc) + vy * (xc / rc);
        }
    }

    /** [B][3][NN] (host, block major) <-> [3][B][NN] (device, field major) */
    __global__ void permute_state(const double* __restrict__ src, double* __restrict__ dst, int B, int NN, size_t FS, int to_device)
    {
        size_t n = size_t(B) * 3 * NN;
        for (size_t k = blockIdx.x * size_t(blockDim.x) + threadIdx.x; k < n; k += size_t(gridDim.x) * blockDim.x)
        {
            size_t cell = k % NN, q = (k / NN) % 3, b = k / (size_t(3) * NN);
            size_t field_major = q * FS + b * NN + cell;
            if (to_device) dst[field_major] = src[k]; else dst[k] = src[field_major];
        }
    }

    __global__ void axpby_kernel(const double* __restrict__ a, double wa, const double* __restrict__ b, double wb, double* __restrict__ dst, size_t n)
    {
        for (size_t k = blockIdx.x * size_t(blockDim.x) + threadIdx.x; k < n; k += size_t(gridDim.x) * blockDim.x)
        {
            dst[k] = a[k] * wa + b[k] * wb;
        }
    }

    template<typename T>
    T* device_upload(const std::vector<T>& v)
    {
        T* p = nullptr;
        M3B_CUDA(cudaMalloc(&p, std::max<size_t>(1, v.size()) * sizeof(T)));
        if (! v.empty()) M3B_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        return p;
    }
}




// ===========================================================================
// device_field_t
// ===========================================================================
device_field_t::device_field_t(std::size_t num_doubles, int device) : count(num_doubles), device(device)
{
    M3B_CUDA(cudaSetDevice(device));
    M3B_CUDA(cudaMalloc(&data, std::max<size_t>(1, count) * sizeof(double)));
}

device_field_t::~device_field_t()
{
    if (data) cudaFree(data);
}




// ===========================================================================
// device_solver_t
// ===========================================================================
device_solver_t::device_solver_t(const solver_data_t& sd, int device, bool general_only, bool tiled_kernel) : impl(new impl_t), device_id(device), force_general(general_only)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        throw std::runtime_error("mara3_b200: no CUDA device is available (this library has no CPU fallback)");
    M3B_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    M3B_CUDA(cudaGetDeviceProperties(&prop, device));
    impl->sm_count = prop.multiProcessorCount;

    N = sd.block_size;
    B = sd.num_local;               // blocks stored on this rank: owned, then ghosts
    BO = sd.num_owned;
    cells = sd.num_local_cells();
    const auto& tree = *sd.tree;
    const auto& part = sd.partition;
    impl->num_global_blocks = sd.num_blocks;
    auto local = [&part] (int global) { return global < 0 ? -1 : part.global_to_local[global]; };

    cudaStream_t s;
    M3B_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    stream_ = own_stream = s;

    // ---- static mesh data
    auto spacing = std::vector<double>(B);
    auto nbr = std::vector<face_nbr_dev_t>(size_t(B) * 4);
    auto nbr9 = std::vector<int>(size_t(B) * 9, -1);
    auto in_gradient_set = std::vector<char>(B, 0);

    for (int b = 0; b < B; ++b)
    {
        const int g = sd.global_block(b);
        spacing[b] = sd.spacing(tree.index(g).level);

        // face tables also for ghost blocks: the any-tree kernels compute gradients of layer-1 ghosts from
        // layer-2 ghosts and walk from a fine ghost back to the coarse block it borders (ids of blocks this
        // rank does not store are -1 and never followed)
        for (int side = 0; side < 4; ++side)
        {
            auto fn = tree.face_neighbor(g, side);
            auto& d = nbr[size_t(b) * 4 + side];
            d.kind = int(fn.kind);
            for (int q = 0; q < 4; ++q) d.leaf[q] = local(fn.leaf[q]);
            d.bx = fn.bx; d.by = fn.by; d.pad = 0;
            for (int q = 0; q < 4; ++q) d.gs[q] = -1;
        }
        if (b >= BO) continue;          // ghost blocks are only read from
        bool regular = true;
        for (int di = -1; di <= 1; ++di)
            for (int dj = -1; dj <= 1; ++dj)
            {
                int l = local(tree.same_level_neighbor(g, di, dj));
                nbr9[size_t(b) * 9 + (di + 1) * 3 + (dj + 1)] = l;
                if (l < 0) regular = false;
            }
        (regular ? impl->regular : impl->irregular).push_back(b);
    }
    // tile shape of the fused kernel; blocks whose size no tile divides all take the general path
    if      (N % 32 == 0) { impl->tile_x = 16; impl->tile_y = 32; impl->strip = ! tiled_kernel; }
    else if (N % 24 == 0) { impl->tile_x = 12; impl->tile_y = 24; }
    else if (N % 16 == 0) { impl->tile_x = 16; impl->tile_y = 16; }
    else if (N % 8 == 0)  { impl->tile_x = 8;  impl->tile_y = 8; }
    else
    {
        impl->irregular.insert(impl->irregular.end(), impl->regular.begin(), impl->regular.end());
        std::sort(impl->irregular.begin(), impl->irregular.end());
        impl->regular.clear();
    }
    // interior blocks (every neighbour owned by this rank) first: they can be updated while the guard zones
    // of the boundary blocks are still in flight
    std::stable_partition(impl->regular.begin(), impl->regular.end(), [&] (int b)
    {
        for (int k = 0; k < 9; ++k) if (nbr9[size_t(b) * 9 + k] >= BO) return false;
        return true;
    });
    impl->num_interior = 0;
    for (int b : impl->regular)
    {
        bool interior = true;
        for (int k = 0; k < 9; ++k) if (nbr9[size_t(b) * 9 + k] >= BO) interior = false;
        impl->num_interior += interior;
    }
    num_regular = int(impl->regular.size());

    auto gslot = std::vector<int>(B, -1);
    auto mark = [&] (int l) { if (l >= 0) in_gradient_set[l] = 1; };

    // the general path reads gradients of its own blocks and of every face neighbour they fetch from
    for (int b : impl->irregular)
    {
        mark(b);
        for (int side = 0; side < 4; ++side) for (int q = 0; q < 4; ++q) mark(nbr[size_t(b) * 4 + side].leaf[q]);
    }
    if (force_general)
        for (int b = 0; b < BO; ++b)        // every owned block takes the any-tree kernels: its face neighbours' gradients too (ghosts included)
        {
            mark(b);
            for (int side = 0; side < 4; ++side) for (int q = 0; q < 4; ++q) mark(nbr[size_t(b) * 4 + side].leaf[q]);
        }
    for (int b = 0; b < B; ++b) if (in_gradient_set[b]) { gslot[b] = int(impl->gradient_blocks.size()); impl->gradient_blocks.push_back(b); }
    for (auto& d : nbr) for (int q = 0; q < 4; ++q) d.gs[q] = d.leaf[q] >= 0 ? gslot[d.leaf[q]] : -1;

    impl->mesh.B = B;
    impl->mesh.N = N;
    impl->mesh.FS = cells;
    impl->mesh.GS = impl->gradient_blocks.size() * size_t(N) * N;
    impl->mesh.xv = device_upload(sd.xv);
    impl->mesh.yv = device_upload(sd.yv);
    impl->mesh.spacing = device_upload(spacing);
    {
        std::vector<double> inv(spacing.size());
        for (size_t k = 0; k < inv.size(); ++k) inv[k] = 1.0 / spacing[k];
        impl->mesh.inv_spacing = device_upload(inv);
    }
    impl->mesh.nbr = device_upload(nbr);
    impl->mesh.nbr9 = device_upload(nbr9);
    impl->mesh.gslot = device_upload(gslot);
    impl->mesh.U0 = device_upload(sd.initial_conserved_u);
    impl->mesh.br = device_upload(sd.buffer_rate_field);
    std::vector<unsigned char> tile_flags_host;
    if (impl->tile_x)
    {
        // static per-tile flags: bit 0 = the buffer-zone rate is non-zero somewhere in the tile
        const int tx = impl->tile_x, ty = impl->tile_y, tiles_y = N / ty, tpb = (N / tx) * tiles_y;
        auto flags = std::vector<unsigned char>(size_t(B) * tpb, 0);
        for (int b = 0; b < B; ++b)
            for (int t = 0; t < tpb; ++t)
            {
                bool any = false;
                for (int i = (t / tiles_y) * tx; i < (t / tiles_y + 1) * tx && ! any; ++i)
                    for (int j = (t % tiles_y) * ty; j < (t % tiles_y + 1) * ty; ++j)
                        if (sd.buffer_rate_field[(size_t(b) * N + i) * N + j] != 0.0) { any = true; break; }
                flags[size_t(b) * tpb + t] = any;
            }
        impl->d_tile_flags = device_upload(flags);
        tile_flags_host = flags;
    }
    impl->d_regular = device_upload(impl->regular);
    auto regular_tile_info = std::vector<tile_info_t>();
    if (impl->tile_x)
    {
        const int tpb = (N / impl->tile_x) * (N / impl->tile_y);
        auto info = std::vector<tile_info_t>(impl->regular.size() * tpb);
        for (size_t r = 0; r < impl->regular.size(); ++r)
            for (int t = 0; t < tpb; ++t)
            {
                auto& ti = info[r * tpb + t];
                ti.b = impl->regular[r];
                for (int k = 0; k < 9; ++k) ti.n9[k] = nbr9[size_t(ti.b) * 9 + k];
                ti.flags = tile_flags_host[size_t(ti.b) * tpb + t] | (t << TILE_POS_SHIFT);
                ti.row = int(r) * tpb + t;
            }
        regular_tile_info = info;       // (uploaded below, with the jump blocks' tiles that touch same-level leaves only behind them)
    }
    // (conserved_q with general_only keeps the any-tree kernels: the reference implementation of that variable set on the device)
    impl->jump_strip = N % 32 == 0 && ! tiled_kernel && (sd.conserve_linear_p || ! general_only);
    if (const char* e = std::getenv("M3B_JUMP_STRIP")) impl->jump_strip = impl->jump_strip && std::atoi(e) != 0;
    if (const char* e = std::getenv("M3B_SERIAL_JUMP")) impl->serial_jump = std::atoi(e) != 0;
    if (const char* e = std::getenv("M3B_JUMP_AFTER")) impl->jump_after_regular = std::atoi(e) != 0;
    if (const char* e = std::getenv("M3B_JUMP_MODE0")) impl->jump_mode0 = std::atoi(e) != 0;
    if (impl->jump_strip)
    {
        // the any-tree list (blocks at refinement jumps, or every owned block with general_only) in 16 x 32 strip tiles
        const int tiles_y = N / 32, tpb = (N / 16) * tiles_y;
        auto list = std::vector<int>();
        if (general_only) for (int b = 0; b < BO; ++b) list.push_back(b);
        else list = impl->irregular;
        // A tile of such a block whose 20 x 36 region (tile + two guard layers) lies in same-level leaves only -- its own block
        // and the <= 3 neighbours it touches -- sees exactly what a tile of a regular block sees: guard cells are copies, no
        // prolongation / restriction, no corrected face (mesh_tree_operators.hpp:223-252, scheme.cpp:614-720 touch the block's
        // sides at the jump only).  With a block of 64^2 in 4 x 2 tiles that is half to three quarters of the tiles of a block
        // with one side at a jump: they are appended to the persistent regular kernel's list (stage_tma) and leave
        // stage_strip<.., JUMP>'s, keeping their rows among the block's.  M3B_SPLIT_JUMP=0: every tile of these blocks through JUMP.
        bool split = ! general_only && sd.conserve_linear_p;        // (stage_tma: linear-momentum variables)
        if (const char* e = std::getenv("M3B_STAGE")) split = split && std::string(e) != "strip";
        if (const char* e = std::getenv("M3B_SPLIT_JUMP")) split = split && std::atoi(e) != 0;
        auto info = std::vector<tile_info_t>();
        for (size_t r = 0; r < list.size(); ++r)
            for (int t = 0; t < tpb; ++t)
            {
                tile_info_t ti;
                const int i0 = (t / tiles_y) * 16, j0 = (t % tiles_y) * 32;
                ti.b = list[r];
                ti.row = int(r) * tpb + t;
                ti.flags = tile_flags_host[size_t(ti.b) * tpb + t] | (t << TILE_POS_SHIFT);
                bool same_level_only = split;
                for (int di = -1; di <= 1 && same_level_only; ++di)
                    for (int dj = -1; dj <= 1; ++dj)
                    {
                        const bool touched = (di == 0 || (di < 0 ? i0 == 0 : i0 + 16 == N)) && (dj == 0 || (dj < 0 ? j0 == 0 : j0 + 32 == N));
                        if (touched && (di | dj) && nbr9[size_t(ti.b) * 9 + (di + 1) * 3 + (dj + 1)] < 0) { same_level_only = false; break; }
                    }
                if (same_level_only)
                {
                    for (int k = 0; k < 9; ++k) ti.n9[k] = nbr9[size_t(ti.b) * 9 + k];     // (-1 where there is no such leaf: never touched)
                    ti.flags |= TILE_JUMP_ROWS;
                    regular_tile_info.push_back(ti);
                    ++impl->num_extra_tiles;
                    continue;
                }
                for (int k = 0; k < 9; ++k) ti.n9[k] = ti.b;        // cells beyond the block come through resolve_cell
                if (i0 == 0       && nbr[size_t(ti.b) * 4 + 0].kind == 2) ti.flags |= 2;
                if (i0 + 16 == N  && nbr[size_t(ti.b) * 4 + 1].kind == 2) ti.flags |= 4;
                if (j0 == 0       && nbr[size_t(ti.b) * 4 + 2].kind == 2) ti.flags |= 8;
                if (j0 + 32 == N  && nbr[size_t(ti.b) * 4 + 3].kind == 2) ti.flags |= 16;
                info.push_back(ti);
            }
        impl->num_jump_tiles = int(info.size());
        if (info.empty()) info.push_back(tile_info_t());
        impl->d_jump_tile_info = device_upload(info);
    }
    if (! regular_tile_info.empty()) impl->d_tile_info = device_upload(regular_tile_info);
    impl->d_irregular = device_upload(impl->irregular);
    impl->d_gradient_blocks = device_upload(impl->gradient_blocks);

    auto all_blocks = std::vector<int>(BO);
    for (int b = 0; b < BO; ++b) all_blocks[b] = b;
    impl->owned.push_back(device_upload(all_blocks));

    impl->model.softening_radius2   = sd.softening_radius * sd.softening_radius;
    impl->model.sink_rate           = sd.sink_rate;
    impl->model.sink_inv_2s2        = 1.0 / (sd.sink_radius * sd.sink_radius) / 2.0;
    impl->model.inv_mach2           = 1.0 / sd.mach_number / sd.mach_number;
    impl->model.inv_mach            = 1.0 / sd.mach_number;
    impl->model.alpha               = sd.alpha;
    impl->model.nu                  = sd.nu;
    impl->model.alpha_cutoff_radius = sd.alpha_cutoff_radius;
    impl->model.density_floor       = sd.density_floor;
    impl->model.axisymmetric_cs2    = sd.axisymmetric_cs2;
    impl->model.qmode               = ! sd.conserve_linear_p;
    impl->model.domain_radius       = sd.domain_radius;
    impl->model.gst_suppr_radius2   = sd.gst_suppr_radius * sd.gst_suppr_radius;
    impl->mesh.qmode                = ! sd.conserve_linear_p;

    // ---- scratch
    size_t max_rows = size_t(BO) * std::max(1, (N / std::max(1, impl->tile_x)) * (N / std::max(1, impl->tile_y))) + B;
    M3B_CUDA(cudaMalloc(&impl->d_partials, max_rows * ROW * sizeof(double)));
    M3B_CUDA(cudaMalloc(&impl->d_partials2, max_rows * ROW * sizeof(double)));
    for (auto& p : impl->d_block_rows) M3B_CUDA(cudaMalloc(&p, size_t(BO + 1) * ROW * sizeof(double)));
    // tile of the tiled any-tree kernels: 16 x 16 where it divides the block, else 12 x 12 (the reference's default block size 24),
    // else 8 x 8; block sizes none of them divides keep one CTA per block
    impl->gtile = N % 16 == 0 ? 16 : (N % 12 == 0 ? 12 : (N % 8 == 0 ? 8 : 0));
    if (impl->gtile)
    {
        const int g = impl->gtile;
        for (auto& p : impl->d_general_tile_rows) M3B_CUDA(cudaMalloc(&p, std::max<size_t>(1, size_t(BO) * (N / g) * (N / g)) * ROW * sizeof(double)));
        M3B_CUDA(cudaFuncSetAttribute(general_update_tiled<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(tile_t<16, 16>))));
        M3B_CUDA(cudaFuncSetAttribute(general_update_tiled<12, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(tile_t<12, 12>))));
        M3B_CUDA(cudaFuncSetAttribute(general_update_tiled<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(tile_t<8, 8>))));
    }
    if (const char* e = std::getenv("M3B_UNTILED_GENERAL")) impl->untiled_general = std::atoi(e) != 0;
    if (const char* e = std::getenv("M3B_MULTI_CTA_FINISH")) impl->multi_cta_finish = std::atoi(e) != 0;
    impl->cta_rows_stride = size_t(BO / FINISH_ROWS_PER_CTA + 1) * ROW;
    M3B_CUDA(cudaMalloc(&impl->d_cta_rows, 2 * impl->cta_rows_stride * sizeof(double)));
    // (default priority: with the highest one the first stage's finish_stage displaces CTAs of the second stage kernel, +2 us)
    M3B_CUDA(cudaStreamCreateWithFlags(&impl->finish_stream, cudaStreamNonBlocking));
    M3B_CUDA(cudaEventCreateWithFlags(&impl->stage_done, cudaEventDisableTiming));
    {
        // above the compute stream's priority: ring gradients -> jump blocks is the critical path of a stage on a nested tree,
        // the regular blocks' CTAs fill the slots it leaves (M3B_JUMP_PRIORITY=0: same priority)
        int least = 0, greatest = 0;
        M3B_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        const char* e = std::getenv("M3B_JUMP_PRIORITY");
        M3B_CUDA(cudaStreamCreateWithPriority(&impl->jump_stream, cudaStreamNonBlocking, e && std::atoi(e) == 0 ? least : greatest));
    }
    M3B_CUDA(cudaEventCreateWithFlags(&impl->gradients_done, cudaEventDisableTiming));
    M3B_CUDA(cudaEventCreateWithFlags(&impl->jump_done, cudaEventDisableTiming));
    M3B_CUDA(cudaEventCreateWithFlags(&impl->side_finish_done, cudaEventDisableTiming));
    M3B_CUDA(cudaEventCreateWithFlags(&impl->fast_prepare_done, cudaEventDisableTiming));
    for (auto& e : impl->positions_done) M3B_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    M3B_CUDA(cudaMalloc(&impl->d_counters, 2 * sizeof(int)));
    M3B_CUDA(cudaMemset(impl->d_counters, 0, 2 * sizeof(int)));
    impl->partial_rows = max_rows;
    M3B_CUDA(cudaMalloc(&impl->d_stage, num_slots * sizeof(stage_t)));
    M3B_CUDA(cudaMallocHost(&impl->h_stage_ring, stage_ring_size * sizeof(stage_t)));
    M3B_CUDA(cudaMalloc(&impl->d_results_local, num_slots * sizeof(stage_result_t)));
    M3B_CUDA(cudaMemset(impl->d_results_local, 0, num_slots * sizeof(stage_result_t)));
    for (auto& e : impl->step_done) M3B_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    {
        // highest priority: the exchange's small kernels must not queue behind the interior update's CTAs
        int least = 0, greatest = 0;
        M3B_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        M3B_CUDA(cudaStreamCreateWithPriority(&impl->comm_stream, cudaStreamNonBlocking, greatest));
    }
    // Measured on 4096^2 (profiles/): +4 % at 4 GPUs but -27 % at 8 GPUs, where the NCCL kernel spins on SMs until
    // the slowest peer arrives and the boundary launch adds a partial wave; off unless asked for.
    if (const char* e = std::getenv("M3B_TRACE")) impl->trace = std::atoi(e) != 0;
    if (const char* e = std::getenv("M3B_OVERLAP_EXCHANGE")) impl->overlap_exchange = std::atoi(e) != 0;
    M3B_CUDA(cudaEventCreateWithFlags(&impl->input_ready, cudaEventDisableTiming));
    M3B_CUDA(cudaEventCreateWithFlags(&impl->halo_ready, cudaEventDisableTiming));
    M3B_CUDA(cudaMalloc(&impl->d_staging, 3 * sd.num_owned_cells() * sizeof(double)));
    M3B_CUDA(cudaMalloc(&impl->d_fail, num_slots * sizeof(fail_dev_t)));
    // stage results live in mapped pinned host memory: finish_stage writes them straight to the host
    M3B_CUDA(cudaHostAlloc(&host_results, num_slots * sizeof(stage_result_t), cudaHostAllocMapped));
    M3B_CUDA(cudaHostGetDevicePointer(&impl->d_results, host_results, 0));
    M3B_CUDA(cudaMemset(impl->d_fail, 0, num_slots * sizeof(fail_dev_t)));
    M3B_CUDA(cudaMalloc(&impl->d_gradients, std::max<size_t>(1, 6 * impl->mesh.GS) * sizeof(double)));
    M3B_CUDA(cudaMalloc(&impl->d_exchange_clock, 6 * sizeof(unsigned long long)));
    M3B_CUDA(cudaMemset(impl->d_exchange_clock, 0, 6 * sizeof(unsigned long long)));
    M3B_CUDA(cudaMalloc(&impl->d_fused_counters, 8 * sizeof(int)));
    M3B_CUDA(cudaMemset(impl->d_fused_counters, 0, 8 * sizeof(int)));
    if (const char* e = std::getenv("M3B_FUSED_EXCHANGE")) impl->fused_exchange = std::atoi(e) != 0;

    // ---- multi-GPU exchange plan (partition.hpp): per peer, the strips in the order both sides agree on
    if (part.is_distributed())
    {
        auto build = [&] (const std::vector<std::vector<halo_region_t>>& lists, std::vector<size_t>& counts, std::vector<size_t>& starts)
        {
            auto entries = std::vector<halo_entry_dev_t>();
            size_t offset = 0;
            counts.assign(part.nranks, 0);
            starts.assign(part.nranks, 0);
            for (int p = 0; p < part.nranks; ++p)
            {
                starts[p] = offset;
                for (const auto& r : lists[p])
                {
                    halo_entry_dev_t e;
                    e.block = r.block;
                    e.i0 = r.di < 0 ? N - 2 : 0; e.ni = r.di ? 2 : N;
                    e.j0 = r.dj < 0 ? N - 2 : 0; e.nj = r.dj ? 2 : N;
                    e.pad = p;              // the other side's rank
                    e.offset = offset;
                    offset += size_t(3) * e.ni * e.nj;
                    entries.push_back(e);
                }
                counts[p] = offset - starts[p];
            }
            return std::make_pair(entries, offset);
        };
        auto send_starts = std::vector<size_t>(), recv_starts = std::vector<size_t>();
        auto [send_entries, send_total] = build(part.send, impl->send_count, send_starts);
        auto [recv_entries, recv_total] = build(part.recv, impl->recv_count, recv_starts);
        impl->num_send_entries = int(send_entries.size());
        impl->num_recv_entries = int(recv_entries.size());
        impl->d_send_entries = device_upload(send_entries);
        impl->d_recv_entries = device_upload(recv_entries);
        impl->send_entries_host = send_entries;
        impl->send_starts_host = send_starts;
        impl->recv_starts_host = recv_starts;
        impl->recv_total = recv_total;
        M3B_CUDA(cudaMalloc(&impl->d_send_buffer, std::max<size_t>(1, send_total) * sizeof(double)));
        M3B_CUDA(cudaMalloc(&impl->d_recv_buffer, std::max<size_t>(1, recv_total) * sizeof(double)));
        for (int p = 0; p < part.nranks; ++p)
        {
            impl->send_ptr.push_back(impl->d_send_buffer + send_starts[p]);
            impl->recv_ptr.push_back(impl->d_recv_buffer + recv_starts[p]);
        }
        impl->halo_bytes_per_exchange = send_total * sizeof(double);
        M3B_CUDA(cudaMalloc(&impl->d_results_all, size_t(part.nranks) * num_slots * sizeof(stage_result_t)));
        M3B_CUDA(cudaMallocHost(&impl->h_results_all, size_t(part.nranks) * num_slots * sizeof(stage_result_t)));
        std::memset(impl->h_results_all, 0, size_t(part.nranks) * num_slots * sizeof(stage_result_t));
    }
    num_ranks = part.nranks;
    rank_ = part.rank;

    auto set_smem = [&] (auto kernel, size_t bytes)
    {
        impl->fused_smem = bytes;
        M3B_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
    };
    if (impl->tile_x == 16 && impl->tile_y == 32) set_smem(stage_fused<16, 32>, sizeof(tile_t<16, 32>));
    if (impl->strip)
    {
        // (3 CTAs x 168 registers and 2 x 220 were measured slower than 4 x 128: DESIGN.md section 3)
        const char* pa = std::getenv("M3B_PREFETCH_AHEAD");
        impl->mesh.prefetch_ahead = pa ? std::atoi(pa) : impl->sm_count * 4;
        set_smem(stage_strip<4, 0, false, 0>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, true, 0>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, false, 0>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 0>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 1>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 2>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, false, 0, false, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, true, 0, false, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, false, 0, false, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 0, false, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 1, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 2, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, false, 0, true, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, true, 0, true, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, false, 0, true, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 0, true, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, false, 0, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 0, true, 0, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, false, 0, true>, sizeof(strip_smem_t));
        set_smem(stage_strip<4, 64, true, 0, true>, sizeof(strip_smem_t));
        impl->fast_eos = ! sd.axisymmetric_cs2 && sd.nu == 0.0 && sd.alpha_cutoff_radius == 0.0 && sd.density_floor == 0.0;
        // stage_tma: linear-momentum variables only (conserved_q keeps stage_strip<.., QMODE>)
        impl->tma = sd.conserve_linear_p;
        if (const char* e = std::getenv("M3B_STAGE")) impl->tma = impl->tma && std::string(e) != "strip";
        if (const char* e = std::getenv("M3B_TMA_CTAS")) impl->tma_ctas_per_sm = std::max(1, std::min(4, std::atoi(e)));
        impl->tma_fast = impl->fast_eos && sd.alpha > 0.0;
        stage_tma_configure();      // stage_tma.cu
    }
    if (impl->tile_x == 12 && impl->tile_y == 24) set_smem(stage_fused<12, 24>, sizeof(tile_t<12, 24>));
    if (impl->tile_x == 16 && impl->tile_y == 16) set_smem(stage_fused<16, 16>, sizeof(tile_t<16, 16>));
    if (impl->tile_x == 8  && impl->tile_y == 8)  set_smem(stage_fused<8, 8>, sizeof(tile_t<8, 8>));
}

device_solver_t::~device_solver_t()
{
    cudaSetDevice(device_id);
    cudaStreamSynchronize(cudaStream_t(stream_));
    for (auto p : {(void*) impl->mesh.xv, (void*) impl->mesh.yv, (void*) impl->mesh.spacing, (void*) impl->mesh.inv_spacing, (void*) impl->mesh.nbr,
                   (void*) impl->mesh.nbr9, (void*) impl->mesh.gslot, (void*) impl->mesh.U0, (void*) impl->mesh.br,
                   (void*) impl->d_regular, (void*) impl->d_irregular, (void*) impl->d_gradient_blocks, (void*) impl->d_gradients,
                   (void*) impl->d_partials, (void*) impl->d_staging, (void*) impl->d_tile_flags, (void*) impl->d_tile_info, (void*) impl->d_jump_tile_info, (void*) impl->d_fail})
        if (p) cudaFree(p);
    for (auto p : impl->owned) cudaFree(p);
    for (auto p : {(void*) impl->d_send_entries, (void*) impl->d_recv_entries, (void*) impl->d_send_buffer, (void*) impl->d_recv_buffer,
                   (void*) impl->d_results_local, (void*) impl->d_results_all})
        if (p) cudaFree(p);
    if (impl->trace && impl->trace_steps.size() > 20)
    {
        const size_t n = impl->trace_names.size(), first = 10;
        auto acc = std::vector<double>(n, 0.0);
        size_t used = 0;
        for (size_t k = first; k < impl->trace_steps.size(); ++k)
        {
            auto& ev = impl->trace_steps[k];
            if (ev.size() != n) continue;
            ++used;
            for (size_t m = 1; m < n; ++m) { float ms = 0; cudaEventElapsedTime(&ms, ev[m - 1], ev[m]); acc[m] += ms; }
            if (k + 1 < impl->trace_steps.size() && impl->trace_steps[k + 1].size() == n)
            { float ms = 0; cudaEventElapsedTime(&ms, ev[n - 1], impl->trace_steps[k + 1][0]); acc[0] += ms; }
        }
        std::fprintf(stderr, "[m3b trace rank %d] %zu steps, us per step:\n  %-28s %8.1f\n", rank_, used, "(gap to next step)", acc[0] / used * 1e3);
        for (size_t m = 1; m < n; ++m) std::fprintf(stderr, "  -> %-25s %8.1f\n", impl->trace_names[m], acc[m] / used * 1e3);
    }
    for (auto p : impl->peer_mailbox) if (p) cudaIpcCloseMemHandle(p);
    for (auto p : {(void*) impl->mailbox, (void*) impl->d_push_entries, (void*) impl->d_push_ticket, (void*) impl->d_ready, (void*) impl->d_exchange_clock, (void*) impl->d_fused_counters}) if (p) cudaFree(p);
    if (impl->h_results_all) cudaFreeHost(impl->h_results_all);
    if (impl->h_stage_ring) cudaFreeHost(impl->h_stage_ring);
    for (auto p : {(void*) impl->d_stage, (void*) impl->d_partials2, (void*) impl->d_block_rows[0], (void*) impl->d_block_rows[1],
                   (void*) impl->d_cta_rows, (void*) impl->d_counters, (void*) impl->d_general_tile_rows[0], (void*) impl->d_general_tile_rows[1]}) if (p) cudaFree(p);
    for (auto e : impl->step_done) if (e) cudaEventDestroy(e);
    if (impl->stage_done) cudaEventDestroy(impl->stage_done);
    if (impl->side_finish_done) cudaEventDestroy(impl->side_finish_done);
    if (impl->fast_prepare_done) cudaEventDestroy(impl->fast_prepare_done);
    for (auto e : impl->positions_done) if (e) cudaEventDestroy(e);
    if (impl->finish_stream) cudaStreamDestroy(impl->finish_stream);
    if (impl->input_ready) cudaEventDestroy(impl->input_ready);
    if (impl->halo_ready) cudaEventDestroy(impl->halo_ready);
    if (impl->comm_stream) cudaStreamDestroy(impl->comm_stream);
    for (auto& ev : impl->timing_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (auto e : impl->event_pool) cudaEventDestroy(e);
    if (host_results) cudaFreeHost(host_results);
    cudaStreamDestroy(cudaStream_t(own_stream));
}

void device_solver_t::set_stream(void* cuda_stream)
{
    sync();
    stream_ = cuda_stream ? cuda_stream : own_stream;
}

void device_solver_t::upload(const double* host, device_field_t& dst)
{
    auto s = cudaStream_t(stream_);
    M3B_CUDA(cudaSetDevice(device_id));
    M3B_CUDA(cudaMemcpyAsync(impl->d_staging, host, 3 * size_t(BO) * N * N * sizeof(double), cudaMemcpyHostToDevice, s));
    permute_state<<<impl->sm_count * 4, 256, 0, s>>>(impl->d_staging, dst.data, BO, N * N, cells, 1);
    ++launches;
    M3B_CUDA(cudaGetLastError());
}

void device_solver_t::download(const device_field_t& src, double* host)
{
    auto s = cudaStream_t(stream_);
    M3B_CUDA(cudaSetDevice(device_id));
    permute_state<<<impl->sm_count * 4, 256, 0, s>>>(src.data, impl->d_staging, BO, N * N, cells, 0);
    ++launches;
    M3B_CUDA(cudaGetLastError());
    M3B_CUDA(cudaMemcpyAsync(host, impl->d_staging, 3 * size_t(BO) * N * N * sizeof(double), cudaMemcpyDeviceToHost, s));
    M3B_CUDA(cudaStreamSynchronize(s));
}




/** [block][3][N][N] of every rank's owned blocks on rank 0 (the layout of download()). */
void device_solver_t::gather_state(const device_field_t& src, double* host_all)
{
    M3B_CUDA(cudaSetDevice(device_id));
    permute_state<<<impl->sm_count * 4, 256, 0, cudaStream_t(stream_)>>>(src.data, impl->d_staging, BO, N * N, cells, 0);
    ++launches;
    gather_blocks(impl->d_staging, size_t(3) * N * N, host_all);
}

void device_solver_t::gather_diagnostic_fields(const device_field_t& src, double* host_all)
{
    M3B_CUDA(cudaSetDevice(device_id));
    diagnostic_fields_kernel<<<impl->sm_count * 4, 256, 0, cudaStream_t(stream_)>>>(impl->mesh, src.data, impl->d_staging, BO);
    ++launches;
    gather_blocks(impl->d_staging, size_t(3) * N * N, host_all);
}

void device_solver_t::disk_totals(const device_field_t& src, double out[2])
{
    M3B_CUDA(cudaSetDevice(device_id));
    auto s = cudaStream_t(stream_);
    double* d = impl->d_staging;                    // [BO][2], idle between transfers
    disk_totals_kernel<<<BO, THREADS, 0, s>>>(impl->mesh, src.data, d);
    ++launches;
    auto h = std::vector<double>(size_t(BO) * 2);
    M3B_CUDA(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    M3B_CUDA(cudaStreamSynchronize(s));
    out[0] = out[1] = 0.0;
    for (int b = 0; b < BO; ++b) { out[0] += h[2 * b]; out[1] += h[2 * b + 1]; }    // tree order, as the reference's .sum()
    if (num_ranks > 1)
    {
        // per-rank sums folded in rank order on every rank
        for (int k = 0; k < 2; ++k)
        {
            auto parts = all_gather_scalar(out[k]);
            out[k] = 0.0;
            for (double v : parts) out[k] += v;
        }
    }
}

void device_solver_t::diagnostic_fields(const device_field_t& src, double* host)
{
    M3B_CUDA(cudaSetDevice(device_id));
    auto s = cudaStream_t(stream_);
    diagnostic_fields_kernel<<<impl->sm_count * 4, 256, 0, s>>>(impl->mesh, src.data, impl->d_staging, BO);
    ++launches;
    M3B_CUDA(cudaMemcpyAsync(host, impl->d_staging, size_t(BO) * 3 * N * N * sizeof(double), cudaMemcpyDeviceToHost, s));
    M3B_CUDA(cudaStreamSynchronize(s));
}

void device_solver_t::copy(const device_field_t& src, device_field_t& dst)
{
    M3B_CUDA(cudaSetDevice(device_id));
    M3B_CUDA(cudaMemcpyAsync(dst.data, src.data, 3 * cells * sizeof(double), cudaMemcpyDeviceToDevice, cudaStream_t(stream_)));
}

void device_solver_t::load_initial(device_field_t& dst)
{
    M3B_CUDA(cudaSetDevice(device_id));
    M3B_CUDA(cudaMemcpyAsync(dst.data, impl->mesh.U0, 3 * cells * sizeof(double), cudaMemcpyDeviceToDevice, cudaStream_t(stream_)));
}

void device_solver_t::combine(const device_field_t& a, double wa, const device_field_t& b, double wb, device_field_t& dst)
{
    M3B_CUDA(cudaSetDevice(device_id));
    axpby_kernel<<<impl->sm_count * 4, 256, 0, cudaStream_t(stream_)>>>(a.data, wa, b.data, wb, dst.data, 3 * cells);
    ++launches;
    M3B_CUDA(cudaGetLastError());
}

void device_solver_t::set_stage_timing(bool on)
{
    if (on && impl->event_pool.size() < 512)
    {
        for (int k = 0; k < 512; ++k) { cudaEvent_t e; M3B_CUDA(cudaEventCreate(&e)); impl->event_pool.push_back(e); }
    }
    stage_timing = on;
}

void device_solver_t::collect_stage_timing()
{
    sync();
    for (auto& ev : impl->timing_events)
    {
        float ms = 0.f;
        M3B_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
        stage_ms_total += ms;
        ++stage_timed_launches;
        impl->event_pool.push_back(ev.first);
        impl->event_pool.push_back(ev.second);
    }
    impl->timing_events.clear();
    auto fold = [&] (std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& list, double& total_us, std::uint64_t& count)
    {
        for (auto& ev : list)
        {
            float ms = 0.f;
            M3B_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
            total_us += ms * 1e3;
            ++count;
            impl->event_pool.push_back(ev.first);
            impl->event_pool.push_back(ev.second);
        }
        list.clear();
    };
    fold(impl->exchange_events, exchange_us_total, exchanges_timed);
    std::uint64_t gaps = 0;
    fold(impl->gap_events, exposed_wait_us_total, gaps);
    unsigned long long words[6] = {0, 0, 0, 0, 0, 0};
    M3B_CUDA(cudaMemcpy(words, impl->d_exchange_clock, sizeof(words), cudaMemcpyDeviceToHost));
    M3B_CUDA(cudaMemset(impl->d_exchange_clock, 0, sizeof(words)));
    result_wait_us_total += words[0] * 1e-3;
    result_waits_timed += words[1];
    unpack_cta_wait_us_total += words[2] * 1e-3;
    unpack_cta_waits += words[3];
    // fused exchange: kernel start -> all strips stored in the neighbours' landing buffers and the flags raised
    exchange_us_total += words[4] * 1e-3;
    exchanges_timed += words[5];
}

void device_solver_t::upload_stage(const stage_inputs_t& inputs, int slot)
{
    auto s = cudaStream_t(stream_);
    stage_t& st = impl->h_stage_ring[impl->ring_next];      // the ring is far longer than the launches in flight
    impl->ring_next = (impl->ring_next + 1) % stage_ring_size;
    st.time = inputs.time;
    st.dt = inputs.dt;
    st.theta = inputs.theta;
    st.x1 = inputs.bodies.body1.x; st.y1 = inputs.bodies.body1.y; st.m1 = inputs.bodies.body1.mass;
    st.x2 = inputs.bodies.body2.x; st.y2 = inputs.bodies.body2.y; st.m2 = inputs.bodies.body2.mass;
    st.vx1 = inputs.bodies.body1.vx; st.vy1 = inputs.bodies.body1.vy;
    st.vx2 = inputs.bodies.body2.vx; st.vy2 = inputs.bodies.body2.vy;
    st.rk_b0 = inputs.rk_b0;
    st.combine = inputs.combine;
    st.compute_dt = inputs.compute_dt;
    M3B_CUDA(cudaMemcpyAsync(impl->d_stage + slot, &st, sizeof(stage_t), cudaMemcpyHostToDevice, s));
}

/** The stage kernels + finish_stage for the inputs already in d_stage[slot].  With `exchange` the guard
 *  zones of `in` are refreshed from the other ranks first, overlapped with the update of the interior blocks. */
void device_solver_t::launch_stage_kernels(const device_field_t& in, const device_field_t* un, device_field_t& out, int slot, bool exchange, int finish_mode, int stage_mode)
{
    auto s = cudaStream_t(stream_);
    if (in.data == out.data) throw std::invalid_argument("launch_stage: in-place stages are not supported");

    const stage_t* st = impl->d_stage + slot;
    double* partials = (slot & 1) ? impl->d_partials2 : impl->d_partials;
    double* block_rows = impl->d_block_rows[slot & 1];
    const double* un_data = un ? un->data : nullptr;
    int num_fused = force_general ? 0 : int(impl->regular.size());
    int num_general = force_general ? BO : int(impl->irregular.size());
    const int* d_general = force_general ? static_cast<const int*>(impl->owned[0]) : impl->d_irregular;
    const int tpb = impl->tile_x ? (N / impl->tile_x) * (N / impl->tile_y) : 0;
    int fused_ctas = num_fused * tpb;
    // the any-tree path in 16 x 16 tiles where the block size allows (one row per tile), else one CTA and one row per block
    const bool jump_strip = impl->jump_strip && impl->strip && ! impl->untiled_general;
    const int gt = impl->untiled_general ? 0 : impl->gtile;
    const int ggtpb = gt ? (N / gt) * (N / gt) : 1;     // tiles of general_gradients_tiled
    const int gtpb = jump_strip ? tpb : ggtpb;
    double* general_rows = gtpb > 1 ? impl->d_general_tile_rows[slot & 1] : block_rows;
    exchange = exchange && num_ranks > 1;
    impl->mark(s, "stage begin");
    bool waiting_tiles = false;
    // stage_tma is persistent: its CTAs hold every register of the SM until the tile list is done, so a tile that waited inside
    // the kernel for the guard-zone unpack could keep the unpack kernel from ever becoming resident.  Its launch is split
    // instead: interior blocks beside the exchange (which runs on its own, higher-priority stream), then the blocks with
    // ghost neighbours once the unpack has finished.  No kernel of this path waits for another kernel of the same GPU.
    const bool tma_exchange = exchange && impl->strip && impl->tma && impl->peer_transport && num_general == 0 && num_fused > 0;
    // (default) the stage kernel does the exchange itself: push first, interior blocks, unpack by the CTAs that reach the blocks
    // with ghost neighbours first -- one launch, nothing on another stream
    const bool fused = tma_exchange && impl->fused_exchange;
    fused_exchange_t X = fused_exchange_t();
    if (fused)
    {
        const unsigned long long counter = ++impl->exchange_counter;
        X.enabled = 1;
        X.push = impl->d_push_entries; X.n_push = impl->num_send_entries;
        X.recv = impl->d_recv_entries; X.n_recv = impl->num_recv_entries;
        X.peers = impl->peers; X.parity = int(counter & 1); X.me = rank_; X.dest_mask = impl->dest_mask; X.counter = counter;
        X.counters = impl->d_fused_counters; X.cset = int(++impl->fused_launches & 1); X.U = const_cast<double*>(in.data);
        X.clock_words = stage_timing ? impl->d_exchange_clock : nullptr;
        exchange = false;
    }
    const bool split_launch = tma_exchange && ! fused && impl->num_recv_entries > 0 && impl->num_interior > 0;
    if (exchange && ! impl->overlap_exchange && ! split_launch)
    {
        waiting_tiles = impl->peer_transport && impl->in_kernel_wait && impl->strip && ! impl->tma && num_general == 0 && impl->num_recv_entries > 0;
        impl->defer_unpack = waiting_tiles;
        exchange_on(stream_, const_cast<device_field_t&>(in));     // plain ordering: exchange, then every block
        impl->defer_unpack = false;
        exchange = false;
        impl->mark(s, "exchange done");
    }

    // blocks [first, first + count) of the regular list: tile rows, block rows and tickets are indexed by list position;
    // extra_tiles: the jump blocks' tiles that touch same-level leaves only (they follow the regular blocks' tiles in d_tile_info,
    // so they can only ride with a launch that ends with the last regular block)
    const int num_extra_tiles = (impl->strip && impl->tma && jump_strip && num_general > 0) ? impl->num_extra_tiles : 0;
    auto launch_fused = [&] (int first, int count, int extra_tiles = 0)
    {
        if (count <= 0 && extra_tiles <= 0) return;
        if (extra_tiles > 0 && first + count != num_fused) throw std::logic_error("launch_fused: the extra tiles follow the last regular block");
        const int ctas = count * tpb + extra_tiles;
        const int* list = impl->d_regular + first;
        double* tiles = partials + size_t(first) * tpb * ROW;
        #define M3B_LAUNCH_FUSED(TX, TY) stage_fused<TX, TY><<<ctas, THREADS, sizeof(tile_t<TX, TY>), s>>>( \
            impl->mesh, impl->model, st, list, in.data, un_data, out.data, tiles, impl->d_fail + slot)
        if (impl->strip && impl->tma)
        {
            stage_tma_launch_t a;
            a.mesh = impl->mesh;
            a.mesh.first_wait_cta = (waiting_tiles || fused) ? std::max(0, impl->num_interior - first) * tpb : 0x7fffffff;
            a.exchange = X;
            a.mesh.ready_flag = impl->d_ready;
            a.mesh.ready_value = impl->exchange_counter;
            a.model = impl->model; a.stage = st; a.tile_info = impl->d_tile_info + size_t(first) * tpb; a.num_tiles = ctas;
            a.Uin = in.data; a.Un = un_data; a.Uout = out.data; a.fail = impl->d_fail + slot;
            a.partials = partials; a.jump_partials = general_rows;      // (rows are addressed through tile_info_t::row)
            a.N = N; a.fast = impl->tma_fast; a.stage_mode = stage_mode;
            // 3 CTAs per SM with two tile buffers unless M3B_TMA_CTAS=4 asks for one buffer and 4: measured on 4096^2, a tile costs a
            // 4-per-SM CTA 1.36 x what it costs a 3-per-SM CTA (10.6 against 7.8 us, its load is exposed), so 589 against 578 us
            // on one GPU; on 2 GPUs 0.680 against 0.634 ms per step (the CTAs also spend 19 instead of 5 us in the fused unpack).
            const int per_sm = impl->tma_ctas_per_sm ? impl->tma_ctas_per_sm : 3;
            a.ctas_per_sm = per_sm;
            a.grid = std::min(ctas, impl->sm_count * per_sm);       // persistent: every CTA walks the tile list with stride `grid`
            stage_tma_launch(a, s);
        }
        else if (impl->strip)
        {
            auto kernel = N == 64 ? (impl->fast_eos ? stage_strip<4, 64, true, 0> : stage_strip<4, 64, false, 0>)
                                  : (impl->fast_eos ? stage_strip<4, 0, true, 0> : stage_strip<4, 0, false, 0>);
            if (N == 64 && impl->fast_eos && stage_mode == 1) kernel = stage_strip<4, 64, true, 1>;
            if (N == 64 && impl->fast_eos && stage_mode == 2) kernel = stage_strip<4, 64, true, 2>;
            if (impl->mesh.qmode)       // conserved_q = (sigma, Sr, Lz): advance_q (scheme.cpp:906-1020) through the same kernel
                kernel = N == 64 ? (impl->fast_eos ? stage_strip<4, 64, true, 0, false, true> : stage_strip<4, 64, false, 0, false, true>)
                                 : (impl->fast_eos ? stage_strip<4, 0, true, 0, false, true> : stage_strip<4, 0, false, 0, false, true>);
            mesh_dev_t mesh = impl->mesh;
            mesh.first_wait_cta = waiting_tiles ? std::max(0, impl->num_interior - first) * tpb : 0x7fffffff;
            mesh.ready_flag = impl->d_ready;
            mesh.ready_value = impl->exchange_counter;
            kernel<<<ctas, STRIP_THREADS, sizeof(strip_smem_t), s>>>(mesh, impl->model, st, impl->d_tile_info + size_t(first) * tpb,
                in.data, un_data, out.data, partials, impl->d_fail + slot, nullptr);
        }
        else if (impl->tile_x == 16 && impl->tile_y == 32) M3B_LAUNCH_FUSED(16, 32);
        else if (impl->tile_x == 12 && impl->tile_y == 24) M3B_LAUNCH_FUSED(12, 24);
        else if (impl->tile_x == 16 && impl->tile_y == 16) M3B_LAUNCH_FUSED(16, 16);
        else                                               M3B_LAUNCH_FUSED(8, 8);
        #undef M3B_LAUNCH_FUSED
        ++launches;
        M3B_CUDA(cudaGetLastError());
    };

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (stage_timing && (fused_ctas > 0 || num_general > 0))
    {
        // events come from a pool: creating them here would delay the launches behind this one
        if (impl->event_pool.size() < 2)
        {
            for (int k = 0; k < 64; ++k) { cudaEvent_t e; M3B_CUDA(cudaEventCreate(&e)); impl->event_pool.push_back(e); }
        }
        e0 = impl->event_pool.back(); impl->event_pool.pop_back();
        e1 = impl->event_pool.back(); impl->event_pool.pop_back();
    }
    // Blocks at refinement jumps through the strip kernel: their launch runs on its own stream beside the regular blocks'
    // (two partial waves of CTAs fill each other's tails); the gradients both read come first.
    auto launch_jump_strip = [&] (cudaStream_t js)
    {
        auto kernel = N == 64 ? (impl->fast_eos ? stage_strip<4, 64, true, 0, true> : stage_strip<4, 64, false, 0, true>)
                              : (impl->fast_eos ? stage_strip<4, 0, true, 0, true> : stage_strip<4, 0, false, 0, true>);
        if (N == 64 && impl->fast_eos && stage_mode == 1 && ! impl->jump_mode0) kernel = stage_strip<4, 64, true, 1, true>;
        if (N == 64 && impl->fast_eos && stage_mode == 2 && ! impl->jump_mode0) kernel = stage_strip<4, 64, true, 2, true>;
        if (impl->mesh.qmode)
            kernel = N == 64 ? (impl->fast_eos ? stage_strip<4, 64, true, 0, true, true> : stage_strip<4, 64, false, 0, true, true>)
                             : (impl->fast_eos ? stage_strip<4, 0, true, 0, true, true> : stage_strip<4, 0, false, 0, true, true>);
        mesh_dev_t mesh = impl->mesh;
        mesh.first_wait_cta = 0x7fffffff;
        const int jump_tiles = force_general ? num_general * tpb : impl->num_jump_tiles;
        if (jump_tiles == 0) return;
        kernel<<<jump_tiles, STRIP_THREADS, sizeof(strip_smem_t), js>>>(mesh, impl->model, st, impl->d_jump_tile_info,
            in.data, un_data, out.data, general_rows, impl->d_fail + slot, impl->d_gradients);
        ++launches;
        M3B_CUDA(cudaGetLastError());
    };
    const int ng = int(impl->gradient_blocks.size());
    bool jump_forked = false, jump_pending = false, e0_recorded = false;
    if (jump_strip && num_general > 0 && ! exchange)
    {
        // fork: gradients and the jump blocks' update on the side stream, the regular blocks' update on the compute stream
        const bool fork = (num_fused > 0 || num_extra_tiles > 0) && ! impl->serial_jump;
        cudaStream_t js = fork ? impl->jump_stream : s;
        if (e0) { M3B_CUDA(cudaEventRecord(e0, s)); e0_recorded = true; }      // the timed region covers the jump blocks' kernels
        if (fork)
        {
            M3B_CUDA(cudaEventRecord(impl->gradients_done, s));
            M3B_CUDA(cudaStreamWaitEvent(js, impl->gradients_done, 0));
        }
        general_gradients_ring<<<ng * 4, 128, size_t(12) * (N + 2) * sizeof(double), js>>>(impl->mesh, st, impl->d_gradient_blocks, in.data, impl->d_gradients);
        ++launches;
        if (fork)
        {
            // jump_after_regular: only the ring gradients run beside the persistent regular kernel (whose CTAs hold their SMs
            // until its tile list is done); the tiles at the jumps follow on the compute stream once both are through
            if (impl->jump_after_regular) jump_pending = true; else launch_jump_strip(js);
            M3B_CUDA(cudaEventRecord(impl->jump_done, js));
            jump_forked = true;
        }
    }
    if (exchange)
    {
        // guard zones travel on their own stream while the interior blocks are updated
        auto& field = const_cast<device_field_t&>(in);
        auto pooled = [&] { cudaEvent_t e = impl->event_pool.back(); impl->event_pool.pop_back(); return e; };
        const bool timed = stage_timing && impl->event_pool.size() >= 8;
        M3B_CUDA(cudaEventRecord(impl->input_ready, s));
        M3B_CUDA(cudaStreamWaitEvent(impl->comm_stream, impl->input_ready, 0));
        cudaEvent_t x0 = nullptr, x1 = nullptr, g0 = nullptr, g1 = nullptr;
        if (timed) { x0 = pooled(); x1 = pooled(); g0 = pooled(); g1 = pooled(); M3B_CUDA(cudaEventRecord(x0, impl->comm_stream)); }
        exchange_on(impl->comm_stream, field);
        M3B_CUDA(cudaEventRecord(impl->halo_ready, impl->comm_stream));
        if (timed) M3B_CUDA(cudaEventRecord(x1, impl->comm_stream));
        if (e0 && ! e0_recorded) M3B_CUDA(cudaEventRecord(e0, s));
        launch_fused(0, impl->num_interior);
        if (timed) M3B_CUDA(cudaEventRecord(g0, s));
        M3B_CUDA(cudaStreamWaitEvent(s, impl->halo_ready, 0));
        if (timed) M3B_CUDA(cudaEventRecord(g1, s));
        launch_fused(impl->num_interior, num_fused - impl->num_interior, num_extra_tiles);
        if (timed) { impl->exchange_events.emplace_back(x0, x1); impl->gap_events.emplace_back(g0, g1); }
        if (num_fused == 0) {}      // (general blocks below run after the wait as well)
    }
    else
    {
        if (e0 && ! e0_recorded) M3B_CUDA(cudaEventRecord(e0, s));
        launch_fused(0, num_fused, num_extra_tiles);
    }
    if (jump_forked)
    {
        M3B_CUDA(cudaStreamWaitEvent(s, impl->jump_done, 0));
        if (jump_pending) launch_jump_strip(s);
    }
    else if (num_general > 0 && jump_strip)
    {
        if (exchange)       // (overlapped exchange: the guard zones have only just arrived)
        {
            general_gradients_ring<<<ng * 4, 128, size_t(12) * (N + 2) * sizeof(double), s>>>(impl->mesh, st, impl->d_gradient_blocks, in.data, impl->d_gradients);
            ++launches;
        }
        launch_jump_strip(s);
    }
    else if (num_general > 0)
    {
        // gradients are needed for the general blocks and every block they can fetch from
        #define M3B_GENERAL_TILED(G_) do { \
            general_gradients_tiled<G_, G_><<<ng * ggtpb, THREADS, 0, s>>>(impl->mesh, st, impl->d_gradient_blocks, in.data, impl->d_gradients); \
            general_update_tiled<G_, G_><<<num_general * gtpb, THREADS, sizeof(tile_t<G_, G_>), s>>>(impl->mesh, impl->model, st, d_general, \
                in.data, impl->d_gradients, un_data, out.data, general_rows, impl->d_fail + slot); } while (0)
        if (ggtpb > 1 && gtpb > 1)
        {
            if (gt == 16) M3B_GENERAL_TILED(16); else if (gt == 12) M3B_GENERAL_TILED(12); else M3B_GENERAL_TILED(8);
        }
        else
        {
            general_gradients<<<ng, THREADS, 0, s>>>(impl->mesh, st, impl->d_gradient_blocks, in.data, impl->d_gradients);
            general_update<<<num_general, THREADS, 0, s>>>(impl->mesh, impl->model, st, d_general,
                in.data, impl->d_gradients, un_data, out.data, general_rows, impl->d_fail + slot);
        }
        #undef M3B_GENERAL_TILED
        launches += 2;
        M3B_CUDA(cudaGetLastError());
    }
    if (e0)         // after the join: regular blocks, jump blocks and their gradients
    {
        M3B_CUDA(cudaEventRecord(e1, s));
        impl->timing_events.emplace_back(e0, e1);
    }
    if (waiting_tiles) M3B_CUDA(cudaStreamWaitEvent(s, impl->halo_ready, 0));     // join the exchange stream
    impl->mark(s, "stage kernels done");
    launch_finish(partials, num_fused, tpb, general_rows, gtpb, num_fused + num_general, slot, finish_mode);
    impl->mark(s, "finish done (or forked)");
    M3B_CUDA(cudaGetLastError());
}

/** finish_mode 0: on the compute stream.  1: on the side stream, beside the next stage (its result is only read at the end of
 *  the step).  2: on the compute stream after the side stream's finish, with the next step's stage inputs written by the last CTA. */
void device_solver_t::launch_finish(const double* tile_rows, int num_fused, int tpb, const double* general_rows, int gtpb, int num_rows, int slot, int finish_mode)
{
    auto s = cudaStream_t(stream_);
    int ctas = std::max(1, (num_rows + FINISH_ROWS_PER_CTA - 1) / FINISH_ROWS_PER_CTA);
    double* cta_rows = impl->d_cta_rows + size_t(slot & 1) * impl->cta_rows_stride;
    prepare_args_t prep = prepare_args_t();

    if (finish_mode == 1)
    {
        M3B_CUDA(cudaEventRecord(impl->stage_done, s));
        M3B_CUDA(cudaStreamWaitEvent(impl->finish_stream, impl->stage_done, 0));
        s = impl->finish_stream;
    }
    if (finish_mode == 2)
    {
        M3B_CUDA(cudaStreamWaitEvent(s, impl->side_finish_done, 0));
        prep = impl->pending_prepare;
    }
    // on the critical path (not beside a stage kernel, finish_mode 1, where eight 1024-thread CTAs would wait for SMs to drain)
    if (num_rows <= FINISH_CLUSTER_MAX_ROWS && ! impl->multi_cta_finish && finish_mode != 1)
        finish_stage_cluster<<<FINISH_CLUSTER, FINISH_CLUSTER_THREADS, 0, s>>>(tile_rows, num_fused, tpb, general_rows, gtpb,
            num_rows, impl->d_stage + slot, impl->d_fail + slot, result_target(slot), prep);
    else
        finish_stage<<<ctas, FINISH_THREADS, 0, s>>>(tile_rows, num_fused, tpb, general_rows, gtpb, num_rows, cta_rows,
            impl->d_counters + (slot & 1), impl->d_stage + slot, impl->d_fail + slot, result_target(slot), prep);
    ++launches;
    M3B_CUDA(cudaGetLastError());
    if (finish_mode == 1) M3B_CUDA(cudaEventRecord(impl->side_finish_done, impl->finish_stream));
}

/** Where finish_stage writes: host-mapped memory for synchronous single-rank use, device memory when the
 *  result still has to be folded over ranks or consumed by prepare_next. */
stage_result_t* device_solver_t::result_target(int slot)
{
    bool on_device = num_ranks > 1;     // a single rank publishes straight to the host (finish_stage also prepares the next step)
    return (on_device ? impl->d_results_local : impl->d_results) + slot;
}

void device_solver_t::launch_stage(const device_field_t& in, const device_field_t* un, device_field_t& out, const stage_inputs_t& inputs, int slot)
{
    M3B_CUDA(cudaSetDevice(device_id));
    if (inputs.combine && ! un) throw std::invalid_argument("launch_stage: combine requires the step-start state");
    upload_stage(inputs, slot);
    launch_stage_kernels(in, un, out, slot, /*exchange*/ true);
}

void device_solver_t::launch_step_async(device_field_t& in, device_field_t& scratch, device_field_t& out, int parity,
                                        const elements_t& elements, double cfl_number, double recommended_time_step, double theta, bool fixed_dt)
{
    auto s = cudaStream_t(stream_);
    M3B_CUDA(cudaSetDevice(device_id));
    const int a = first_async_slot + 2 * parity, b = a + 1;
    const int na = first_async_slot + 2 * (1 - parity), nb = na + 1;

    step_config_t cfg;
    cfg.elements = elements;
    cfg.cfl_number = cfl_number;
    cfg.recommended_time_step = recommended_time_step;
    cfg.theta = theta;
    cfg.fixed_dt = fixed_dt;

    const bool split_prepare = num_ranks == 1 || impl->peer_transport;
    if (split_prepare)
    {
        if (impl->fresh_pipeline)
        {
            // the next step's first stage needs its body positions: normally prepared a step ahead, here for the first time
            // (after the host's upload of this step's inputs, which is queued on the compute stream)
            M3B_CUDA(cudaEventRecord(impl->fast_prepare_done, s));
            M3B_CUDA(cudaStreamWaitEvent(impl->finish_stream, impl->fast_prepare_done, 0));
            prepare_positions<<<1, 32, 0, impl->finish_stream>>>(cfg, impl->d_stage + a, nullptr, impl->d_stage + na);
            ++launches;
            M3B_CUDA(cudaEventRecord(impl->positions_done[1 - parity], impl->finish_stream));
        }
        M3B_CUDA(cudaStreamWaitEvent(s, impl->positions_done[parity], 0));         // positions of this step's first stage
    }
    impl->fresh_pipeline = false;

    if (num_ranks == 1)
    {
        // the first stage's rows are folded on the side stream while the second stage runs; the second
        // stage's finish_stage also writes time and dt of the next step's stages (no separate prepare_next)
        impl->pending_prepare.enabled = 1;
        impl->pending_prepare.cfg = cfg;
        impl->pending_prepare.current_a = impl->d_stage + a;
        impl->pending_prepare.next_a = impl->d_stage + na;
        impl->pending_prepare.next_b = impl->d_stage + nb;
        launch_stage_kernels(in, nullptr, scratch, a, /*exchange*/ true, 1, 1);
        M3B_CUDA(cudaStreamWaitEvent(s, impl->positions_done[1 - parity], 0));     // positions of the second stage
        launch_stage_kernels(scratch, &in, out, b, /*exchange*/ true, 2, fixed_dt ? 0 : 2);
    }
    else if (impl->peer_transport)
    {
        // as on one rank, the first stage's rows are folded beside the second stage; the per-rank results then
        // travel through the mailboxes and prepare_next_peer folds them (no NCCL call in the step)
        auto& pp = impl->pending_prepare;
        pp.enabled = 2;
        pp.cfg = cfg;
        pp.current_a = impl->d_stage + a;
        pp.next_a = impl->d_stage + na;
        pp.next_b = impl->d_stage + nb;
        pp.peers = impl->peers;
        pp.local = impl->d_results_local;
        pp.host_results = impl->d_results;
        pp.me = rank_; pp.nranks = num_ranks; pp.slot_stride = num_slots; pp.slot_a = a; pp.slot_b = b;
        pp.counter = ++impl->step_counter;
        pp.clock_words = stage_timing ? impl->d_exchange_clock : nullptr;
        launch_stage_kernels(in, nullptr, scratch, a, /*exchange*/ true, 1, 1);
        M3B_CUDA(cudaStreamWaitEvent(s, impl->positions_done[1 - parity], 0));
        launch_stage_kernels(scratch, &in, out, b, /*exchange*/ true, 2, fixed_dt ? 0 : 2);
    }
    else
    {
        launch_stage_kernels(in, nullptr, scratch, a, /*exchange*/ true, 0, 1);
        launch_stage_kernels(scratch, &in, out, b, /*exchange*/ true, 0, fixed_dt ? 0 : 2);
        const size_t doubles = num_slots * sizeof(stage_result_t) / sizeof(double);
        impl->comm->all_gather(reinterpret_cast<const double*>(impl->d_results_local), reinterpret_cast<double*>(impl->d_results_all), doubles, stream_);
        prepare_next<<<1, 32, 0, s>>>(impl->d_results_all, num_ranks, num_slots, a, b, cfg, impl->d_stage, impl->d_stage + na, impl->d_stage + nb, impl->d_results);
        ++launches;
        M3B_CUDA(cudaGetLastError());
    }
    if (split_prepare)
    {
        // the slow half: positions of the next step's second stage and of the first stage of the step after it
        M3B_CUDA(cudaEventRecord(impl->fast_prepare_done, s));
        M3B_CUDA(cudaStreamWaitEvent(impl->finish_stream, impl->fast_prepare_done, 0));
        prepare_positions<<<1, 32, 0, impl->finish_stream>>>(cfg, impl->d_stage + na, impl->d_stage + nb, impl->d_stage + a);
        ++launches;
        M3B_CUDA(cudaEventRecord(impl->positions_done[parity], impl->finish_stream));
        M3B_CUDA(cudaGetLastError());
    }
    impl->mark(s, "prepare done");
    impl->end_step();
    M3B_CUDA(cudaEventRecord(impl->step_done[parity], s));
}

void device_solver_t::upload_step_inputs(int parity, const stage_inputs_t& first, const stage_inputs_t& second)
{
    M3B_CUDA(cudaSetDevice(device_id));
    upload_stage(first, first_async_slot + 2 * parity);
    upload_stage(second, first_async_slot + 2 * parity + 1);
    impl->fresh_pipeline = true;
}

void device_solver_t::wait_step(int parity)
{
    M3B_CUDA(cudaEventSynchronize(impl->step_done[parity]));
}

stage_result_t device_solver_t::async_result(int parity, int stage) const
{
    const auto& r = host_results[first_async_slot + 2 * parity + stage];     // already folded over ranks by prepare_next
    if (num_ranks > 1 && impl->peer_transport) throw_if_called_off(r.pad);
    return r;
}

int device_solver_t::async_slot(int parity, int stage) const
{
    return first_async_slot + 2 * parity + stage;
}

void device_solver_t::launch_max_timestep(const device_field_t& in, double time, const two_body_t& bodies, int slot)
{
    auto s = cudaStream_t(stream_);
    M3B_CUDA(cudaSetDevice(device_id));
    auto inputs = stage_inputs_t();
    inputs.time = time;
    inputs.dt = 0.0;
    inputs.theta = 0.0;
    inputs.bodies = bodies;
    upload_stage(inputs, slot);
    double* block_rows = impl->d_block_rows[slot & 1];
    max_timestep_kernel<<<BO, THREADS, 0, s>>>(impl->mesh, impl->model, impl->d_stage + slot, in.data, block_rows);
    ++launches;
    launch_finish(nullptr, 0, 1, block_rows, 1, BO, slot, 0);
}








/** Some rank gave up waiting for a peer (bounded_wait_sys): name both in an exception, which the C ABI turns into M3B_ERROR. */
void m3b::throw_if_called_off(unsigned int word)
{
    if (word == 0) return;
    const unsigned waiting = (word - 1) & 255u, awaited = (word - 1) >> 8;
    throw std::runtime_error("mara3_b200: rank " + std::to_string(waiting) + " waited longer than the deadline (M3B_SPIN_DEADLINE_MS) for rank "
        + std::to_string(awaited) + " to deliver its guard zones or stage results; the run was called off on every rank");
}

stage_result_t device_solver_t::stage_result(int slot) const
{
    if (num_ranks == 1) return host_results[slot];
    if (impl->peer_transport) throw_if_called_off(host_results[slot].pad);

    // fold the ranks in rank order: every rank computes the same bits
    auto r = stage_result_t();
    r.dt_min = 1e300;
    for (int p = 0; p < num_ranks; ++p)
    {
        const auto& q = impl->h_results_all[size_t(p) * num_slots + slot];
        for (int k = 0; k < 16; ++k) r.sums[k] += q.sums[k];
        r.work[0] += q.work[0];
        r.work[1] += q.work[1];
        r.dt_min = std::min(r.dt_min, q.dt_min);
        r.num_negative += q.num_negative;
    }
    return r;
}

unsigned int device_solver_t::local_num_negative(int slot) const
{
    return num_ranks == 1 ? host_results[slot].num_negative : impl->h_results_all[size_t(rank_) * num_slots + slot].num_negative;
}

void device_solver_t::sync()
{
    M3B_CUDA(cudaSetDevice(device_id));
    M3B_CUDA(cudaStreamSynchronize(cudaStream_t(stream_)));
    if (impl->finish_stream) M3B_CUDA(cudaStreamSynchronize(impl->finish_stream));
    if (impl->comm_stream) M3B_CUDA(cudaStreamSynchronize(impl->comm_stream));
}

std::vector<offender_t> device_solver_t::offenders(int slot)
{
    fail_dev_t f;
    M3B_CUDA(cudaSetDevice(device_id));
    M3B_CUDA(cudaStreamSynchronize(cudaStream_t(stream_)));
    M3B_CUDA(cudaMemcpy(&f, impl->d_fail + slot, sizeof(fail_dev_t), cudaMemcpyDeviceToHost));
    auto n = std::min<unsigned>(f.pad, max_offenders);
    auto v = std::vector<offender_t>(f.list, f.list + n);
    std::sort(v.begin(), v.end(), [] (const offender_t& a, const offender_t& b) { return a.block != b.block ? a.block < b.block : a.cell < b.cell; });
    return v;
}
