/**
 * any_tree_kernels.cuh -- the kernels behind the persistent / strip stage kernels: the generic tiled kernel for regular blocks
 * whose size is not a multiple of 32 (stage_fused<TX, TY>) and the any-tree path for blocks at refinement jumps when the strip
 * kernel's JUMP variant does not apply (general_gradients*, general_update*; also the on-device reference of that variant), with
 * the CTA-wide reductions they share.  Included by kernels.cu only (one translation unit with the launch logic).
 *
 * Replaces, per RK stage, the reference's phases P1-P8 and P11 of binary::advance_u (Mara3 src/subprog_binary_scheme.cpp:790-904):
 *
 *   stage_fused<TX,TY>   for "regular" blocks (all 8 neighbours are same-level leaves): one CTA per TX x TY tile; the tile plus
 *                        a 2-cell halo is read once from HBM, primitives / PLM differences / face fluxes live in shared memory,
 *                        and the updated cells are written once.
 *   general_*            for blocks touching a refinement jump: guard values are fetched through per-face neighbour tables
 *                        with the reference's prolongation (injection) / restriction (2x2 mean) rules for primitives AND
 *                        gradients (mesh_tree_operators.hpp:223-252), and coarse faces next to finer blocks take the sum of
 *                        the two fine fluxes (scheme.cpp:614-720).
 */
#pragma once
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
