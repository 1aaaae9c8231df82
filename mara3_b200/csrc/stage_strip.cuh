/**
 * stage_strip.cuh -- the fused RK-stage kernel for regular blocks whose size is a multiple of 32.
 *
 * Same update as stage_fused (kernels.cu) -- phases P1-P8 + P11 of binary::advance_u (Mara3
 * src/subprog_binary_scheme.cpp:790-904) -- organised around the warp:
 *
 *   CTA  = 4 warps on a 16 x 32 tile; lane <-> column j (the contiguous direction), warp w owns the
 *          strip of rows 4w .. 4w+3.
 *   P0   tile + 2-cell halo: coalesced row loads (15-30 in flight per thread), conserved -> primitive,
 *        into shared memory.
 *   P1   PLM differences on tile + 1 halo, marching down the strip with the x-stencil in registers.
 *   P2/3 one rolled loop down the strip: each iteration computes the low-x and low-y HLLE + viscous
 *        fluxes of a row, issues the loads the row's update will need, and finishes the update of
 *        the row before (its high-x flux is the x-face just computed, its high-y flux comes from
 *        lane + 1 by shuffle); tile-boundary faces and strip-to-strip fluxes go through two small
 *        shared arrays.  Cells are written once; 16 running sums, the CFL minimum and the
 *        negative-density count are folded per CTA.
 *
 * The loop is rolled on purpose: fully unrolled the kernel is 110 KB of SASS and a fifth of all
 * issue slots stall on instruction fetch (profiles/); rolled it stays inside the instruction cache.
 */
#pragma once

namespace
{
    constexpr int SX = 16, SY = 32, STRIP = 4, STRIP_THREADS = 128;

    struct strip_smem_t
    {
        double P[3][SX + 4][SY + 4];        // primitives sigma, vx, vy on tile + 2 halo
        double G[6][SX + 2][SY + 2];        // un-divided PLM differences d/dx (3), d/dy (3) on tile + 1 halo
        double XB[3][5][SY];                // x-face fluxes at rows 4, 8, 12 (strip starts) and 16 (tile boundary)
        double YB[3][SX];                   // y-face fluxes at the tile's high-y boundary
        double xv[SX + 1];
        double yv[SY + 1];
        double red[STRIP_THREADS / 32][NUM_SUMS + 1];
    };

    __device__ __forceinline__ double shfl_down1(double v)
    {
        return __shfl_down_sync(0xffffffffu, v, 1);
    }

    /** x-face between tile cells (li - 1, lj) and (li, lj), 0 <= li <= SX */
    __device__ __forceinline__ void strip_x_face(const strip_smem_t& T, const model_t& model, const stage_t& S, double inv_h, int li, int lj, double F[3])
    {
        eos_t e = eos_at_face(model, S, T.xv[li], 0.5 * (T.yv[lj] + T.yv[lj + 1]));
        prim_t pl = {T.P[0][li + 1][lj + 2], T.P[1][li + 1][lj + 2], T.P[2][li + 1][lj + 2]};
        prim_t pr = {T.P[0][li + 2][lj + 2], T.P[1][li + 2][lj + 2], T.P[2][li + 2][lj + 2]};
        prim_t gl = {T.G[0][li][lj + 1], T.G[1][li][lj + 1], T.G[2][li][lj + 1]};
        prim_t gr = {T.G[0][li + 1][lj + 1], T.G[1][li + 1][lj + 1], T.G[2][li + 1][lj + 1]};
        face_flux<0>(e, pl, pr, gl, gr, T.G[4][li][lj + 1], T.G[5][li][lj + 1], T.G[4][li + 1][lj + 1], T.G[5][li + 1][lj + 1], 0.5, inv_h, F);
    }

    /** y-face between tile cells (li, lj - 1) and (li, lj), 0 <= lj <= SY */
    __device__ __forceinline__ void strip_y_face(const strip_smem_t& T, const model_t& model, const stage_t& S, double inv_h, int li, int lj, double F[3])
    {
        eos_t e = eos_at_face(model, S, 0.5 * (T.xv[li] + T.xv[li + 1]), T.yv[lj]);
        prim_t pl = {T.P[0][li + 2][lj + 1], T.P[1][li + 2][lj + 1], T.P[2][li + 2][lj + 1]};
        prim_t pr = {T.P[0][li + 2][lj + 2], T.P[1][li + 2][lj + 2], T.P[2][li + 2][lj + 2]};
        prim_t gl = {T.G[3][li + 1][lj], T.G[4][li + 1][lj], T.G[5][li + 1][lj]};
        prim_t gr = {T.G[3][li + 1][lj + 1], T.G[4][li + 1][lj + 1], T.G[5][li + 1][lj + 1]};
        face_flux<1>(e, pl, pr, gl, gr, T.G[1][li + 1][lj], T.G[2][li + 1][lj], T.G[1][li + 1][lj + 1], T.G[2][li + 1][lj + 1], 0.5, inv_h, F);
    }

    __global__ void __launch_bounds__(STRIP_THREADS, 4) stage_strip(
        mesh_dev_t mesh, model_t model, const stage_t S, const int* __restrict__ regular_list,
        const unsigned char* __restrict__ tile_flags,
        const double* __restrict__ Uin, const double* __restrict__ Un, double* __restrict__ Uout,
        double* __restrict__ partials, fail_dev_t* fail)
    {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        strip_smem_t& T = *reinterpret_cast<strip_smem_t*>(smem_raw);

        const int N = mesh.N;
        const int tiles_y = N / SY, tiles_per_block = (N / SX) * tiles_y;
        const int r_index = blockIdx.x / tiles_per_block;
        const int b  = regular_list[r_index];
        const int t  = blockIdx.x % tiles_per_block;
        const int i0 = (t / tiles_y) * SX, j0 = (t % tiles_y) * SY;
        const size_t FS = mesh.FS;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const bool has_buffer = tile_flags[size_t(b) * tiles_per_block + t] & 1;

        // the update phase's inputs are first touched ~10 us from now: pull their lines into L2 already
        if (lane < 2 * STRIP)
        {
            const size_t c = (size_t(b) * N + (i0 + STRIP * warp + (lane >> 1))) * N + j0 + 16 * (lane & 1);
            if (has_buffer)
            {
                asm volatile("prefetch.global.L2 [%0];" :: "l"(mesh.br + c));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(mesh.U0 + c));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(mesh.U0 + FS + c));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(mesh.U0 + 2 * FS + c));
            }
            if (S.combine)
            {
                asm volatile("prefetch.global.L2 [%0];" :: "l"(Un + c));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(Un + FS + c));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(Un + 2 * FS + c));
            }
        }

        // ------------------------------------------------------------------ phase 0: load + primitives
        {
            // region column c <-> global column j0 - 2 + c; group A: c = lane, group B: c = 32 + lane (lane < 4)
            const int gjA = j0 - 2 + lane, gjB = j0 + 30 + lane;
            const int djA = gjA < 0 ? -1 : 0, djB = gjB >= N ? 1 : 0;
            const int colA = gjA - djA * N, colB = gjB - djB * N;
            const int* n9 = mesh.nbr9 + size_t(b) * 9;
            const int nbAm = n9[0 * 3 + djA + 1], nbA0 = n9[1 * 3 + djA + 1], nbAp = n9[2 * 3 + djA + 1];
            const int nbBm = n9[0 * 3 + djB + 1], nbB0 = n9[1 * 3 + djB + 1], nbBp = n9[2 * 3 + djB + 1];
            double uA[5][3], uB[5][3];

            #pragma unroll
            for (int k = 0; k < 5; ++k)
            {
                const int r = warp + 4 * k;                     // region row, 0..19 <-> global row i0 - 2 + r
                const int gi = i0 - 2 + r;
                const int di = gi < 0 ? -1 : (gi >= N ? 1 : 0);
                const int row = gi - di * N;
                const int nbA = di < 0 ? nbAm : (di > 0 ? nbAp : nbA0);
                const int nbB = di < 0 ? nbBm : (di > 0 ? nbBp : nbB0);
                const size_t cA = (size_t(nbA) * N + row) * N + colA;
                const size_t cB = (size_t(nbB) * N + row) * N + colB;
                uA[k][0] = Uin[cA]; uA[k][1] = Uin[FS + cA]; uA[k][2] = Uin[2 * FS + cA];
                if (lane < 4) { uB[k][0] = Uin[cB]; uB[k][1] = Uin[FS + cB]; uB[k][2] = Uin[2 * FS + cB]; }
            }
            #pragma unroll
            for (int k = 0; k < 5; ++k)
            {
                const int r = warp + 4 * k;
                prim_t p = cons_to_prim(uA[k][0], uA[k][1], uA[k][2]);
                T.P[0][r][lane] = p.s; T.P[1][r][lane] = p.vx; T.P[2][r][lane] = p.vy;
                if (lane < 4)
                {
                    prim_t q = cons_to_prim(uB[k][0], uB[k][1], uB[k][2]);
                    T.P[0][r][32 + lane] = q.s; T.P[1][r][32 + lane] = q.vx; T.P[2][r][32 + lane] = q.vy;
                }
            }
            if (warp == 0 && lane <= SX) T.xv[lane] = mesh.xv[size_t(b) * (N + 1) + i0 + lane];
            if (warp == 1) T.yv[lane] = mesh.yv[size_t(b) * (N + 1) + j0 + lane];
            if (warp == 2 && lane == 0) T.yv[SY] = mesh.yv[size_t(b) * (N + 1) + j0 + SY];
        }
        __syncthreads();

        // ------------------------------------------------------------------ phase 1: PLM differences
        {
            // gradient rows g = 0..17 (<-> P row g + 1): warps take 5, 5, 4, 4 consecutive rows;
            // lane <-> gradient column c = lane (<-> P column lane + 1)
            const int g0 = warp < 2 ? 5 * warp : 10 + 4 * (warp - 2);
            const int g1 = g0 + (warp < 2 ? 5 : 4);
            double pm[3], pc[3];
            #pragma unroll
            for (int q = 0; q < 3; ++q) { pm[q] = T.P[q][g0][lane + 1]; pc[q] = T.P[q][g0 + 1][lane + 1]; }

            #pragma unroll 1
            for (int g = g0; g < g1; ++g)
            {
                #pragma unroll
                for (int q = 0; q < 3; ++q)
                {
                    double pp = T.P[q][g + 2][lane + 1];
                    T.G[q][g][lane]     = plm_diff(pm[q], pc[q], pp, S.theta);
                    T.G[3 + q][g][lane] = plm_diff(T.P[q][g + 1][lane], pc[q], T.P[q][g + 1][lane + 2], S.theta);
                    pm[q] = pc[q]; pc[q] = pp;
                }
            }
            // gradient columns 32, 33: 36 cells, taken by the two warps with one row less
            const int k2 = warp == 2 ? lane : (warp == 3 && lane < 4 ? 32 + lane : -1);
            if (k2 >= 0)
            {
                const int g = k2 >> 1, c = 32 + (k2 & 1);
                #pragma unroll
                for (int q = 0; q < 3; ++q)
                {
                    double ctr = T.P[q][g + 1][c + 1];
                    T.G[q][g][c]     = plm_diff(T.P[q][g][c + 1], ctr, T.P[q][g + 2][c + 1], S.theta);
                    T.G[3 + q][g][c] = plm_diff(T.P[q][g + 1][c], ctr, T.P[q][g + 1][c + 2], S.theta);
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------------ phases 2 + 3, one rolled loop
        // iteration r = -1 : the tile-boundary faces (high-x row by warp 0, high-y column by half of warp 1)
        // iteration r = 0..3: the low-x and low-y faces of strip row r; loads for cell r; update of cell r - 1
        // iteration r = 4 : the update of cell 3, whose high-x flux comes from the next strip (or the boundary row)
        const double h = mesh.spacing[b], inv_h = 1.0 / h;
        const int li0 = STRIP * warp, lj = lane;
        const double dt_over_h = S.dt * inv_h;
        const double yc = 0.5 * (T.yv[lj] + T.yv[lj + 1]);
        const size_t c0 = (size_t(b) * N + (i0 + li0)) * N + (j0 + lj);

        double sums[NUM_SUMS];
        #pragma unroll
        for (int k = 0; k < NUM_SUMS; ++k) sums[k] = 0.0;
        double dtmin = 1e300;
        double FxLo[3] = {0, 0, 0}, FyLo[3] = {0, 0, 0};       // low-side fluxes of the cell awaiting its update
        double u[3] = {0, 0, 0}, un[3] = {0, 0, 0}, u0[3] = {0, 0, 0}, br = 0.0;

        #pragma unroll 1
        for (int r = -1; r <= STRIP; ++r)
        {
            double FxNew[3] = {0, 0, 0}, FyNew[3] = {0, 0, 0};
            double v[3] = {0, 0, 0}, vn[3] = {0, 0, 0}, v0[3] = {0, 0, 0}, vbr = 0.0;

            if (r == 1) __syncthreads();        // XB (written at r = -1, 0) and YB (r = -1) are complete

            if (r >= 0 && r < STRIP)
            {
                // inputs of this row's update, consumed one iteration from now
                const size_t c = c0 + size_t(r) * N;
                v[0] = Uin[c]; v[1] = Uin[FS + c]; v[2] = Uin[2 * FS + c];
                if (has_buffer) { vbr = mesh.br[c]; v0[0] = mesh.U0[c]; v0[1] = mesh.U0[FS + c]; v0[2] = mesh.U0[2 * FS + c]; }
                if (S.combine)  { vn[0] = Un[c]; vn[1] = Un[FS + c]; vn[2] = Un[2 * FS + c]; }
            }
            // which faces this thread computes in this iteration
            const bool do_x = r < 0 ? warp == 0 : r < STRIP;
            const bool do_y = r < 0 ? (warp == 1 && lane < SX) : r < STRIP;
            const int xi = r < 0 ? SX : li0 + r, xj = lj;
            const int yi = r < 0 ? lane : li0 + r, yj = r < 0 ? SY : lj;

            if (do_x) strip_x_face(T, model, S, inv_h, xi, xj, FxNew);
            if (do_y) strip_y_face(T, model, S, inv_h, yi, yj, FyNew);

            if (r < 0)
            {
                if (do_x) { T.XB[0][4][lj] = FxNew[0]; T.XB[1][4][lj] = FxNew[1]; T.XB[2][4][lj] = FxNew[2]; }
                if (do_y) { T.YB[0][lane] = FyNew[0]; T.YB[1][lane] = FyNew[1]; T.YB[2][lane] = FyNew[2]; }
            }
            else if (r == 0)
            {
                if (warp > 0) { T.XB[0][warp][lj] = FxNew[0]; T.XB[1][warp][lj] = FxNew[1]; T.XB[2][warp][lj] = FxNew[2]; }
            }
            else
            {
                // update of cell r - 1: its high-x flux is this iteration's x-face, or the next strip's first face
                const int li = li0 + r - 1;
                const size_t c = c0 + size_t(r - 1) * N;
                double hx[3], hy[3];
                #pragma unroll
                for (int q = 0; q < 3; ++q)
                {
                    hx[q] = r < STRIP ? FxNew[q] : T.XB[q][warp + 1][lj];
                    double up = shfl_down1(FyLo[q]);
                    hy[q] = lane == 31 ? T.YB[q][li] : up;
                }
                const double x = 0.5 * (T.xv[li] + T.xv[li + 1]);
                double src[3], y1, y2;
                source_terms(model, S, x, yc, u[0], u[1], u[2], u0[0], u0[1], u0[2], br, src, sums, y1, y2);

                double n0 = u[0] - ((hx[0] - FxLo[0]) + (hy[0] - FyLo[0])) * dt_over_h + src[0];
                double n1 = u[1] - ((hx[1] - FxLo[1]) + (hy[1] - FyLo[1])) * dt_over_h + src[1];
                double n2 = u[2] - ((hx[2] - FxLo[2]) + (hy[2] - FyLo[2])) * dt_over_h + src[2];

                if (n0 < 0.0) report_negative(fail, b, (i0 + li) * N + j0 + lj, n0);

                if (S.combine)
                {
                    const double w = 1.0 - S.rk_b0;
                    n0 = un[0] * S.rk_b0 + n0 * w;
                    n1 = un[1] * S.rk_b0 + n1 * w;
                    n2 = un[2] * S.rk_b0 + n2 * w;
                }
                Uout[c] = n0; Uout[FS + c] = n1; Uout[2 * FS + c] = n2;

                if (S.compute_dt) dtmin = fmin(dtmin, h / max_wavespeed(model, S, x, yc, y1, y2, n0, n1, n2));
            }
            #pragma unroll
            for (int q = 0; q < 3; ++q)
            {
                FxLo[q] = FxNew[q]; FyLo[q] = FyNew[q];
                u[q] = v[q]; un[q] = vn[q]; u0[q] = v0[q];
            }
            br = vbr;
        }

        // ------------------------------------------------------------------ fold the CTA's sums
        // per-warp shuffle tree over the groups that can be non-zero, then 4 warps through shared memory
        const bool sinks_touched = __any_sync(0xffffffffu, sums[ACC_MASS] != 0.0 || sums[ACC_MASS + 1] != 0.0);
        #pragma unroll
        for (int k = 0; k < NUM_SUMS; ++k)
        {
            const bool sink_group = k < GRV_FX, buffer_group = k >= BUF_M;
            double v = 0.0;
            if ((! sink_group || sinks_touched) && (! buffer_group || has_buffer))      // warp-uniform
            {
                v = sums[k];
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            }
            if (lane == 0) T.red[warp][k] = v;
        }
        {
            double m = dtmin;
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) T.red[warp][NUM_SUMS] = m;
        }
        __syncthreads();
        if (threadIdx.x <= NUM_SUMS)
        {
            const int k = threadIdx.x;
            double* row = partials + size_t(blockIdx.x) * ROW;
            double a = T.red[0][k], bq = T.red[1][k], cq = T.red[2][k], d = T.red[3][k];
            row[k] = k == NUM_SUMS ? fmin(fmin(a, bq), fmin(cq, d)) : ((a + bq) + (cq + d)) * (h * h);
        }
    }
}
