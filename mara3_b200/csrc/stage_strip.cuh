/**
 * stage_strip.cuh -- the fused RK-stage kernel for regular blocks whose size is a multiple of 32.
 *
 * Same update as stage_fused (kernels.cu) -- phases P1-P8 + P11 of binary::advance_u (Mara3
 * src/subprog_binary_scheme.cpp:790-904) -- organised around the warp:
 *
 *   CTA  = 4 warps on a 16 x 32 tile; lane <-> column j (the contiguous direction), warp w owns the
 *          strip of rows 4w .. 4w+3.
 *   P0   tile + 2-cell halo as 360 sixteen-byte chunks (9 LDG.128 per thread), conserved -> primitive,
 *        into shared memory.
 *   P1   PLM differences on tile + 1 halo, marching down the strip with the x-stencil in registers.
 *   P2/3 a software-pipelined loop down the strip: each iteration issues the loads for the update of
 *        the row before, computes the low-x and low-y HLLE + viscous fluxes of its row while they are
 *        in flight, and finishes that update (its high-x flux is the x-face just computed, its high-y
 *        flux comes from lane + 1 by shuffle); tile-boundary faces and strip-to-strip fluxes go
 *        through two small shared arrays.  Cells are written once; 16 running sums, the CFL minimum and the
 *        negative-density count are folded per CTA.
 *
 * JUMP = true is the same kernel for blocks at refinement jumps (any 2:1 balanced tree): cells beyond a block side
 * are not recomputed from the neighbour's state but taken as the reference's guard fill delivers them -- primitives
 * through get_cell_block (injection from a coarser block, 2 x 2 mean of finer ones, mesh_tree_operators.hpp:223-252)
 * and PLM gradients from general_gradients' per-block arrays through the same operator (scheme.cpp:810-813) -- and
 * the faces on a side whose neighbour is finer are replaced by the sum of the two fine fluxes (correct_fluxes_*,
 * scheme.cpp:614-720).  Everything inside the block runs exactly as in the regular variant.
 *
 * The strip loop is unrolled in the regular variants (59 KB of SASS, 6 % of the issue slots wait for instructions) and
 * rolled in the JUMP variants, whose extra guard-ring and flux-correction code had pushed instruction fetch to the top
 * stall (profiles/r01_jump_strip_ncu_summary.txt).  QMODE = true evolves conserved_q = (sigma, Sr, Lz) (advance_q,
 * scheme.cpp:906-1020): velocities are recovered at each cell's position in its own block during the tile load, face
 * fluxes are converted with to_angmom_fluxes, the sources are source_terms_q.
 */
#pragma once

namespace
{
    struct strip_smem_t
    {
        double P[3][SX + 4][SY + 4];        // primitives sigma, vx, vy on tile + 2 halo
        double G[6][SX + 2][SY + 2];        // un-divided PLM differences d/dx (3), d/dy (3) on tile + 1 halo
        double XB[3][5][SY];                // x-face fluxes at rows 4, 8, 12 (strip starts) and 16 (tile boundary)
        double YB[3][SX];                   // y-face fluxes at the tile's high-y boundary
        double YLo[3][SX];                  // JUMP: corrected y-face fluxes at a low-y block side with a finer neighbour
        double XLo[3][SY];                  // JUMP: the same at a low-x block side
        double xv[SX + 1];
        double yv[SY + 1];
        // squared-distance tables of the tile's face / centre coordinates to the two bodies (k = 0, 1) and
        // to the origin (k = 2); the y tables of the bodies already hold the softening rs^2
        double x2v[3][SX + 1], x2c[3][SX];      // (xv - x_k)^2, (xc - x_k)^2
        double y2v[3][SY + 1], y2c[3][SY];      // (yv - y_k)^2 [+ rs^2], (yc - y_k)^2 [+ rs^2]
        double red[STRIP_THREADS / 32][NUM_SUMS + 1];
        double sinks[STRIP_THREADS / 32][8];     // per-warp sink sums (ACC_MASS .. ACC_LZ), see source_terms<.., WARP_SINKS>
    };

    template<bool FAST>
    __device__ __forceinline__ eos_t strip_x_eos(const strip_smem_t& T, const model_t& model, const stage_t& S, int li, int lj)
    {
        return eos_from_distances<FAST>(model, S, T.x2v[0][li] + T.y2c[0][lj], T.x2v[1][li] + T.y2c[1][lj], T.x2v[2][li] + T.y2c[2][lj]);
    }

    template<bool FAST>
    __device__ __forceinline__ eos_t strip_y_eos(const strip_smem_t& T, const model_t& model, const stage_t& S, int li, int lj)
    {
        return eos_from_distances<FAST>(model, S, T.x2c[0][li] + T.y2v[0][lj], T.x2c[1][li] + T.y2v[1][lj], T.x2c[2][li] + T.y2v[2][lj]);
    }

    /** x-face between tile cells (li - 1, lj) and (li, lj), 0 <= li <= SX.  QMODE: flux of (sigma, Sr, Lz), to_angmom_fluxes (scheme.cpp:199-214). */
    template<bool QMODE = false>
    __device__ __forceinline__ void strip_x_face(const strip_smem_t& T, const eos_t& e, double inv_h, int li, int lj, double F[3], const model_t* model = nullptr)
    {
        prim_t pl = {T.P[0][li + 1][lj + 2], T.P[1][li + 1][lj + 2], T.P[2][li + 1][lj + 2]};
        prim_t pr = {T.P[0][li + 2][lj + 2], T.P[1][li + 2][lj + 2], T.P[2][li + 2][lj + 2]};
        prim_t gl = {T.G[0][li][lj + 1], T.G[1][li][lj + 1], T.G[2][li][lj + 1]};
        prim_t gr = {T.G[0][li + 1][lj + 1], T.G[1][li + 1][lj + 1], T.G[2][li + 1][lj + 1]};
        face_flux<0>(e, pl, pr, gl, gr, T.G[4][li][lj + 1], T.G[5][li][lj + 1], T.G[4][li + 1][lj + 1], T.G[5][li + 1][lj + 1], 0.5, inv_h, F);
        if (QMODE) to_angmom_fluxes<0>(*model, T.xv[li], 0.5 * (T.yv[lj] + T.yv[lj + 1]), F);
    }

    /** y-face between tile cells (li, lj - 1) and (li, lj), 0 <= lj <= SY */
    template<bool QMODE = false>
    __device__ __forceinline__ void strip_y_face(const strip_smem_t& T, const eos_t& e, double inv_h, int li, int lj, double F[3], const model_t* model = nullptr)
    {
        prim_t pl = {T.P[0][li + 2][lj + 1], T.P[1][li + 2][lj + 1], T.P[2][li + 2][lj + 1]};
        prim_t pr = {T.P[0][li + 2][lj + 2], T.P[1][li + 2][lj + 2], T.P[2][li + 2][lj + 2]};
        prim_t gl = {T.G[3][li + 1][lj], T.G[4][li + 1][lj], T.G[5][li + 1][lj]};
        prim_t gr = {T.G[3][li + 1][lj + 1], T.G[4][li + 1][lj + 1], T.G[5][li + 1][lj + 1]};
        face_flux<1>(e, pl, pr, gl, gr, T.G[1][li + 1][lj], T.G[2][li + 1][lj], T.G[1][li + 1][lj + 1], T.G[2][li + 1][lj + 1], 0.5, inv_h, F);
        if (QMODE) to_angmom_fluxes<1>(*model, 0.5 * (T.xv[li] + T.xv[li + 1]), T.yv[lj], F);
    }

    /**
     * JUMP: flux through the face of tile cell `t` on block side `side` (0 low-x, 1 high-x for AXIS 0; 2 low-y, 3 high-y for
     * AXIS 1) whose neighbour is finer, per unit length of this block: the sum of the two fine faces' fluxes times their
     * lengths as the fine blocks compute them (correct_fluxes_*, scheme.cpp:614-720; block_fluxes_u, :472-516).  The fine
     * blocks' guard cell there is this block's cell by injection (mesh_prolong_restrict.hpp:161-196), primitives and
     * gradients alike, so that side comes from shared memory; the fine cells and their gradients are read directly.
     */
    template<int AXIS, bool QMODE>
    __device__ __forceinline__ void jump_corrected_face(const strip_smem_t& T, const mesh_dev_t& mesh, const model_t& model, const stage_t& S,
        const double* __restrict__ Uin, const double* __restrict__ Gphys, int b, int side, int t, int i0, int j0, int N, double h, double inv_h, double F[3])
    {
        const int4* np = reinterpret_cast<const int4*>(mesh.nbr + size_t(b) * 4 + side);
        const int4 nA = __ldg(np), nB = __ldg(np + 1), nC = __ldg(np + 2);
        const bool lo = (side & 1) == 0;            // the fine blocks lie on this block's low side: they are the left state
        const int near = lo ? 1 : 0;
        const int li = AXIS == 0 ? (lo ? 0 : SX - 1) : t;
        const int lj = AXIS == 0 ? t : (lo ? 0 : SY - 1);
        const int kt = 2 * (AXIS == 0 ? j0 + lj : i0 + li);
        const int hi = kt >= N;
        const int child = AXIS == 0 ? near + 2 * hi : hi + 2 * near;
        const int leaf = child == 0 ? nA.y : (child == 1 ? nA.z : (child == 2 ? nA.w : nB.x));
        const int gs   = child == 0 ? nC.x : (child == 1 ? nC.y : (child == 2 ? nC.z : nC.w));
        const int kf = kt - hi * N, ff = lo ? N - 1 : 0;
        const double* __restrict__ xvf = mesh.xv + size_t(leaf) * (N + 1);
        const double* __restrict__ yvf = mesh.yv + size_t(leaf) * (N + 1);
        const size_t FS = mesh.FS, GS = mesh.GS;

        // all loads of both fine faces first
        double u[2][3], g[2][5], va[2], vb[2];
        const double fixed = AXIS == 0 ? xvf[lo ? N : 0] : yvf[lo ? N : 0];
        #pragma unroll
        for (int m = 0; m < 2; ++m)
        {
            const int fi = AXIS == 0 ? ff : kf + m, fj = AXIS == 0 ? kf + m : ff;
            const size_t c  = (size_t(leaf) * N + fi) * N + fj;
            const size_t gc = (size_t(gs) * N + fi) * N + fj;
            u[m][0] = Uin[c]; u[m][1] = Uin[FS + c]; u[m][2] = Uin[2 * FS + c];
            #pragma unroll
            for (int q = 0; q < 3; ++q) g[m][q] = Gphys[(3 * AXIS + q) * GS + gc];              // longitudinal
            g[m][3] = Gphys[(3 * (1 - AXIS) + 1) * GS + gc];                                     // transverse, vx and vy
            g[m][4] = Gphys[(3 * (1 - AXIS) + 2) * GS + gc];
            va[m] = AXIS == 0 ? yvf[kf + m] : xvf[kf + m];
            vb[m] = AXIS == 0 ? yvf[kf + m + 1] : xvf[kf + m + 1];
        }
        // QMODE: the fine cells' centre coordinate along AXIS (their velocity is recovered at their own position)
        const double fc = QMODE ? 0.5 * ((AXIS == 0 ? xvf[ff] : yvf[ff]) + (AXIS == 0 ? xvf[ff + 1] : yvf[ff + 1])) : 0.0;
        const prim_t pc = {T.P[0][li + 2][lj + 2], T.P[1][li + 2][lj + 2], T.P[2][li + 2][lj + 2]};
        const prim_t gc = {T.G[3 * AXIS][li + 1][lj + 1] * inv_h, T.G[3 * AXIS + 1][li + 1][lj + 1] * inv_h, T.G[3 * AXIS + 2][li + 1][lj + 1] * inv_h};
        const double hcx = T.G[3 * (1 - AXIS) + 1][li + 1][lj + 1] * inv_h, hcy = T.G[3 * (1 - AXIS) + 2][li + 1][lj + 1] * inv_h;

        double acc[3] = {0.0, 0.0, 0.0};
        #pragma unroll
        for (int m = 0; m < 2; ++m)
        {
            prim_t pf = cons_to_prim(u[m][0], u[m][1], u[m][2]);
            const prim_t gf = {g[m][0], g[m][1], g[m][2]};
            const double mid = 0.5 * (va[m] + vb[m]), len = vb[m] - va[m];
            if (QMODE) angmom_to_linear_fast(AXIS == 0 ? fc : mid, AXIS == 0 ? mid : fc, pf.vx, pf.vy, pf.vx, pf.vy);
            const eos_t e = eos_at_face(model, S, AXIS == 0 ? fixed : mid, AXIS == 0 ? mid : fixed);
            double Fm[3];
            if (lo) face_flux<AXIS>(e, pf, pc, gf, gc, g[m][3], g[m][4], hcx, hcy, 0.25 * h, 1.0, Fm);
            else    face_flux<AXIS>(e, pc, pf, gc, gf, hcx, hcy, g[m][3], g[m][4], 0.25 * h, 1.0, Fm);
            if (QMODE) to_angmom_fluxes<AXIS>(model, AXIS == 0 ? fixed : mid, AXIS == 0 ? mid : fixed, Fm);
            if (m == 0) { acc[0] = Fm[0] * len; acc[1] = Fm[1] * len; acc[2] = Fm[2] * len; }
            else        { acc[0] += Fm[0] * len; acc[1] += Fm[1] * len; acc[2] += Fm[2] * len; }
        }
        F[0] = acc[0] * inv_h; F[1] = acc[1] * inv_h; F[2] = acc[2] * inv_h;
    }

    template<int MIN_CTAS, int NB, bool FAST, int MODE, bool JUMP = false, bool QMODE = false>
    __global__ void __launch_bounds__(STRIP_THREADS, MIN_CTAS) stage_strip(
        mesh_dev_t mesh, model_t model, const stage_t* __restrict__ stage_ptr, const tile_info_t* __restrict__ tile_info,
        const double* __restrict__ Uin, const double* __restrict__ Un, double* __restrict__ Uout,
        double* partials, fail_dev_t* fail, const double* __restrict__ Gphys = nullptr)
    {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        strip_smem_t& T = *reinterpret_cast<strip_smem_t*>(smem_raw);

        const stage_t S = *stage_ptr;       // written by the host or by prepare_next of the step before
        // MODE 1 / 2: first / last stage of an RK2 step with adaptive dt, flags known at compile time; 0: read them from S
        const bool combine = MODE == 0 ? S.combine != 0 : MODE == 2, compute_dt = MODE == 0 ? S.compute_dt != 0 : MODE == 2;
        const int N = NB ? NB : mesh.N;       // NB: block size known at compile time (addresses fold into immediates)
        const int tiles_y = N / SY, tiles_per_block = (N / SX) * tiles_y;
        const int4* ti4 = reinterpret_cast<const int4*>(tile_info + blockIdx.x);
        const int4 tiA = __ldg(ti4), tiB = __ldg(ti4 + 1), tiC = __ldg(ti4 + 2);
        const int n9[9] = {tiA.y, tiA.z, tiA.w, tiB.x, tiB.y, tiB.z, tiB.w, tiC.x, tiC.y};
        const int b  = tiA.x;
        const int t  = tiC.z >> TILE_POS_SHIFT;       // position of the tile in its block
        const int i0 = (t / tiles_y) * SX, j0 = (t % tiles_y) * SY;
        const size_t FS = mesh.FS;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const bool has_buffer = tiC.z & 1;
        // JUMP: block sides of this tile whose neighbour is finer (tile_info flags, bits 1-4: low-x, high-x, low-y, high-y)
        const bool finer_lo_x = JUMP && (tiC.z & 2), finer_hi_x = JUMP && (tiC.z & 4);
        const bool finer_lo_y = JUMP && (tiC.z & 8), finer_hi_y = JUMP && (tiC.z & 16);

        // JUMP: the one-cell guard ring beyond the block's sides (no corners: a face only needs its two cells), one thread per
        // ring cell.  The side's neighbour record is requested now; the cells it points at are loaded beside the tile.
        bool ring_on = false;
        int ring_li = 0, ring_lj = 0, ring_ii = 0, ring_jj = 0;
        int4 rnA = make_int4(0, 0, 0, 0), rnB = rnA, rnC = rnA;
        if (JUMP)
        {
            const int k = threadIdx.x;          // 2 (SY + 2) + 2 SX = 100 ring cells
            if      (k < SY + 2)            { ring_li = -1; ring_lj = k - 1; }
            else if (k < 2 * (SY + 2))      { ring_li = SX; ring_lj = k - (SY + 2) - 1; }
            else if (k < 2 * (SY + 2) + SX) { ring_li = k - 2 * (SY + 2); ring_lj = -1; }
            else                            { ring_li = k - 2 * (SY + 2) - SX; ring_lj = SY; }
            const int gi = i0 + ring_li, gj = j0 + ring_lj;
            const bool out_i = gi < 0 || gi >= N, out_j = gj < 0 || gj >= N;
            ring_on = k < 2 * (SY + 2) + 2 * SX && (out_i != out_j);
            if (ring_on)
            {
                const int side = gi < 0 ? 0 : (gi >= N ? 1 : (gj < 0 ? 2 : 3));
                ring_ii = gi < 0 ? N - 1 : (gi >= N ? 0 : gi);
                ring_jj = gj < 0 ? N - 1 : (gj >= N ? 0 : gj);
                const int4* np = reinterpret_cast<const int4*>(mesh.nbr + size_t(b) * 4 + side);
                rnA = __ldg(np); rnB = __ldg(np + 1); rnC = __ldg(np + 2);
            }
        }

        // one "generation" of resident CTAs ahead: pull the tile that a later CTA of this SM slot will load from HBM into L2
        {
            const int ahead = blockIdx.x + mesh.prefetch_ahead;
            if (mesh.prefetch_ahead > 0 && ahead < gridDim.x && threadIdx.x < 96)
            {
                const int rb = __ldg(&tile_info[ahead].b), rt = __ldg(&tile_info[ahead].flags) >> TILE_POS_SHIFT;
                const int f = threadIdx.x >> 5, row = (threadIdx.x & 31) >> 1, half = threadIdx.x & 1;
                const size_t c = (size_t(rb) * N + ((rt / tiles_y) * SX + row)) * N + (rt % tiles_y) * SY + 16 * half;
                asm volatile("prefetch.global.L2 [%0];" :: "l"(Uin + f * FS + c));
            }
        }
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
            if (combine)
            {
                asm volatile("prefetch.global.L2 [%0];" :: "l"(Un + c));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(Un + FS + c));
                asm volatile("prefetch.global.L2 [%0];" :: "l"(Un + 2 * FS + c));
            }
        }

        if (lane < 8) T.sinks[warp][lane] = 0.0;

        // multi-GPU: a tile of a block with ghost neighbours waits until the guard-zone unpack has finished
        if (int(blockIdx.x) >= mesh.first_wait_cta)
        {
            if (threadIdx.x == 0)
            {
                unsigned long long v;
                // (bounded: the unpack kernel that publishes the flag is itself bounded by the peers' deadline, bounded_wait_sys;
                // if it never became resident beside this kernel -- a profiler serialising kernels, a foreign stream -- give up
                // after ~70 s of SM clocks rather than hang: the step then fails its parity, it does not block the GPU)
                const long long t0 = clock64();
                do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mesh.ready_flag) : "memory"); }
                while (v < mesh.ready_value && clock64() - t0 < (1ll << 37));
            }
            __syncthreads();
        }

        // ------------------------------------------------------------------ phase 0: load + primitives
        double ring_grad[6];
        {
            // The (SX + 4) x (SY + 4) region is 20 rows of 18 sixteen-byte chunks (two cells in y); columns j0 - 2 and
            // N are even, so a chunk never straddles two blocks.  Thread <-> (row rr + 7 k, chunk cc), k = 0, 1, 2:
            // 9 LDG.128 per thread, conserved -> primitive for two cells at a time, 9 STS.128.
            const int rr = threadIdx.x / 18, cc = threadIdx.x - 18 * rr;
            const int gj = j0 - 2 + 2 * cc;
            const int dj = gj < 0 ? -1 : (gj >= N ? 1 : 0);
            const long col = gj - dj * N;
            const int nlo  = dj < 0 ? n9[0] : (dj > 0 ? n9[2] : n9[1]);
            const int nmid = dj < 0 ? n9[3] : (dj > 0 ? n9[5] : n9[4]);
            const int nhi  = dj < 0 ? n9[6] : (dj > 0 ? n9[8] : n9[7]);
            const double* __restrict__ U1 = Uin + FS;
            const double* __restrict__ U2 = Uin + 2 * FS;
            double2 uc[3][3];
            double qx[3], qy0[3], qy1[3];           // QMODE: centre coordinates of the chunk's two cells in THEIR block
            bool inside[3] = {true, true, true};    // JUMP: chunks beyond a block side are not loaded here (guard ring below)

            #pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                const int row = rr + 7 * k;
                if (rr < 7 && row < SX + 4)
                {
                    const int gi = i0 - 2 + row;
                    const int di = gi < 0 ? -1 : (gi >= N ? 1 : 0);
                    const int nb = di < 0 ? nlo : (di > 0 ? nhi : nmid);
                    const long c = (long(nb) * N + (gi - di * N)) * N + col;
                    if (JUMP) inside[k] = di == 0 && dj == 0;
                    if (inside[k])
                    {
                        uc[k][0] = *reinterpret_cast<const double2*>(Uin + c);
                        uc[k][1] = *reinterpret_cast<const double2*>(U1 + c);
                        uc[k][2] = *reinterpret_cast<const double2*>(U2 + c);
                    }
                    if (QMODE)
                    {
                        const double* xn = mesh.xv + size_t(nb) * (N + 1) + (gi - di * N);
                        const double* yn = mesh.yv + size_t(nb) * (N + 1) + col;
                        const double ymid = yn[1];
                        qx[k] = 0.5 * (xn[0] + xn[1]); qy0[k] = 0.5 * (yn[0] + ymid); qy1[k] = 0.5 * (ymid + yn[2]);
                    }
                }
            }
            // JUMP: guard ring, second half: get_cell_block (mesh_tree_operators.hpp:223-252) for the ring cell -- the neighbour's
            // own cell, the coarser neighbour's cell by injection, or the mean of the finer neighbour's 2 x 2 cells (axis 0 first,
            // mesh_prolong_restrict.hpp:124-132, 262-272) -- for the primitives and, scaled to this block's un-divided differences,
            // for the neighbour's gradients, which replace what phase 1 computes at the ring cells.
            if (JUMP && ring_on)
            {
                const int kind = rnA.x;
                const size_t GS = mesh.GS;
                const double hb = mesh.spacing[b];
                prim_t p;
                if (kind != 2)
                {
                    const int ci = kind == 0 ? ring_ii : (rnB.y * N + ring_ii) / 2, cj = kind == 0 ? ring_jj : (rnB.z * N + ring_jj) / 2;
                    const size_t c  = (size_t(rnA.y) * N + ci) * N + cj;
                    const size_t gc = (size_t(rnC.x) * N + ci) * N + cj;
                    const double u0 = Uin[c], u1 = Uin[FS + c], u2 = Uin[2 * FS + c];
                    #pragma unroll
                    for (int q = 0; q < 6; ++q) ring_grad[q] = Gphys[q * GS + gc] * hb;
                    p = cons_to_prim(u0, u1, u2);
                    if (QMODE)
                    {
                        const double* xn = mesh.xv + size_t(rnA.y) * (N + 1) + ci;
                        const double* yn = mesh.yv + size_t(rnA.y) * (N + 1) + cj;
                        angmom_to_linear_fast(0.5 * (xn[0] + xn[1]), 0.5 * (yn[0] + yn[1]), p.vx, p.vy, p.vx, p.vy);
                    }
                }
                else
                {
                    // the four fine cells lie in one child: rows fi, fi + 1, columns fj, fj + 1 (fj even: 16-byte loads)
                    const int fi = 2 * ring_ii, fj = 2 * ring_jj;
                    const int child = (fi >= N) + 2 * (fj >= N);
                    const int leaf = child == 0 ? rnA.y : (child == 1 ? rnA.z : (child == 2 ? rnA.w : rnB.x));
                    const int gs   = child == 0 ? rnC.x : (child == 1 ? rnC.y : (child == 2 ? rnC.z : rnC.w));
                    const size_t c  = (size_t(leaf) * N + (fi % N)) * N + (fj % N);
                    const size_t gc = (size_t(gs) * N + (fi % N)) * N + (fj % N);
                    double2 a[3], d[3];
                    #pragma unroll
                    for (int q = 0; q < 3; ++q)
                    {
                        a[q] = *reinterpret_cast<const double2*>(Uin + q * FS + c);
                        d[q] = *reinterpret_cast<const double2*>(Uin + q * FS + c + N);
                    }
                    #pragma unroll
                    for (int q = 0; q < 6; ++q)
                    {
                        const double2 ga = *reinterpret_cast<const double2*>(Gphys + q * GS + gc);
                        const double2 gd = *reinterpret_cast<const double2*>(Gphys + q * GS + gc + N);
                        ring_grad[q] = (((ga.x + gd.x) * 0.5 + (ga.y + gd.y) * 0.5) * 0.5) * hb;
                    }
                    prim_t p00 = cons_to_prim(a[0].x, a[1].x, a[2].x), p01 = cons_to_prim(a[0].y, a[1].y, a[2].y);
                    prim_t p10 = cons_to_prim(d[0].x, d[1].x, d[2].x), p11 = cons_to_prim(d[0].y, d[1].y, d[2].y);
                    if (QMODE)
                    {
                        const double* xn = mesh.xv + size_t(leaf) * (N + 1) + (fi % N);
                        const double* yn = mesh.yv + size_t(leaf) * (N + 1) + (fj % N);
                        const double xm = xn[1], ym = yn[1];
                        const double x0 = 0.5 * (xn[0] + xm), x1 = 0.5 * (xm + xn[2]), y0 = 0.5 * (yn[0] + ym), y1 = 0.5 * (ym + yn[2]);
                        angmom_to_linear_fast(x0, y0, p00.vx, p00.vy, p00.vx, p00.vy);
                        angmom_to_linear_fast(x0, y1, p01.vx, p01.vy, p01.vx, p01.vy);
                        angmom_to_linear_fast(x1, y0, p10.vx, p10.vy, p10.vx, p10.vy);
                        angmom_to_linear_fast(x1, y1, p11.vx, p11.vy, p11.vx, p11.vy);
                    }
                    p = {((p00.s + p10.s) * 0.5 + (p01.s + p11.s) * 0.5) * 0.5,
                         ((p00.vx + p10.vx) * 0.5 + (p01.vx + p11.vx) * 0.5) * 0.5,
                         ((p00.vy + p10.vy) * 0.5 + (p01.vy + p11.vy) * 0.5) * 0.5};
                }
                T.P[0][ring_li + 2][ring_lj + 2] = p.s; T.P[1][ring_li + 2][ring_lj + 2] = p.vx; T.P[2][ring_li + 2][ring_lj + 2] = p.vy;
            }
            // coordinate tables while the loads are in flight
            {
                const double* xvg = mesh.xv + size_t(b) * (N + 1) + i0;
                const double* yvg = mesh.yv + size_t(b) * (N + 1) + j0;
                const double bx[3] = {S.x1, S.x2, 0.0}, by[3] = {S.y1, S.y2, 0.0};
                const double soft[3] = {model.softening_radius2, model.softening_radius2, 0.0};
                if (warp == 0)
                {
                    if (lane <= SX)
                    {
                        const double xv = xvg[lane];
                        T.xv[lane] = xv;
                        #pragma unroll
                        for (int k = 0; k < 3; ++k) T.x2v[k][lane] = (xv - bx[k]) * (xv - bx[k]);
                    }
                    if (lane < SX)
                    {
                        const double xc = 0.5 * (xvg[lane] + xvg[lane + 1]);
                        #pragma unroll
                        for (int k = 0; k < 3; ++k) T.x2c[k][lane] = (xc - bx[k]) * (xc - bx[k]);
                    }
                }
                else if (warp == 1)
                {
                    const double yv = yvg[lane], yc = 0.5 * (yvg[lane] + yvg[lane + 1]);
                    T.yv[lane] = yv;
                    #pragma unroll
                    for (int k = 0; k < 3; ++k)
                    {
                        T.y2v[k][lane] = fma(yv - by[k], yv - by[k], soft[k]);
                        T.y2c[k][lane] = fma(yc - by[k], yc - by[k], soft[k]);
                    }
                }
                else if (warp == 2 && lane == 0)
                {
                    const double yv = yvg[SY];
                    T.yv[SY] = yv;
                    #pragma unroll
                    for (int k = 0; k < 3; ++k) T.y2v[k][SY] = fma(yv - by[k], yv - by[k], soft[k]);
                }
            }
            #pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                const int row = rr + 7 * k;
                if (rr < 7 && row < SX + 4 && inside[k])
                {
                    // iso2d::recover_primitive (physics_iso2d.hpp:351-362) for the chunk's two cells
                    const double ia = fast_rcp(uc[k][0].x), ib = fast_rcp(uc[k][0].y);
                    *reinterpret_cast<double2*>(&T.P[0][row][2 * cc]) = uc[k][0];
                    double2 v1 = make_double2(uc[k][1].x * ia, uc[k][1].y * ib), v2 = make_double2(uc[k][2].x * ia, uc[k][2].y * ib);
                    if (QMODE)
                    {
                        // recover_primitive(Q, x) (physics_iso2d.hpp:376-389): velocity from (Sr, Lz) / sigma at the cell's own position
                        angmom_to_linear_fast(qx[k], qy0[k], v1.x, v2.x, v1.x, v2.x);
                        angmom_to_linear_fast(qx[k], qy1[k], v1.y, v2.y, v1.y, v2.y);
                    }
                    *reinterpret_cast<double2*>(&T.P[1][row][2 * cc]) = v1;
                    *reinterpret_cast<double2*>(&T.P[2][row][2 * cc]) = v2;
                }
            }
        }
        __syncthreads();

        // ------------------------------------------------------------------ phase 1: PLM differences
        {
            // gradient rows g = 0..17 (<-> P row g + 1): warps take 5, 5, 4, 4 consecutive rows;
            // lane <-> gradient column c = lane (<-> P column lane + 1).  Marching down the rows the
            // backward x-difference of a row is the forward difference of the row before.
            const int g0 = warp < 2 ? 5 * warp : 10 + 4 * (warp - 2);
            const int g1 = g0 + (warp < 2 ? 5 : 4);
            double pc[3], dl[3];
            #pragma unroll
            for (int q = 0; q < 3; ++q) { pc[q] = T.P[q][g0 + 1][lane + 1]; dl[q] = pc[q] - T.P[q][g0][lane + 1]; }

            #pragma unroll 2
            for (int g = g0; g < g1; ++g)
            {
                // all nine loads first: behind a store to T.G the compiler will not hoist a load from T.P
                double pp[3], yl[3], yr[3];
                #pragma unroll
                for (int q = 0; q < 3; ++q) { pp[q] = T.P[q][g + 2][lane + 1]; yl[q] = T.P[q][g + 1][lane]; yr[q] = T.P[q][g + 1][lane + 2]; }
                #pragma unroll
                for (int q = 0; q < 3; ++q)
                {
                    const double dr = pp[q] - pc[q];
                    T.G[q][g][lane]     = plm_from_differences(dl[q], dr, S.theta);
                    T.G[3 + q][g][lane] = plm_from_differences(pc[q] - yl[q], yr[q] - pc[q], S.theta);
                    pc[q] = pp[q]; dl[q] = dr;
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
                    const double ctr = T.P[q][g + 1][c + 1];
                    T.G[q][g][c]     = plm_from_differences(ctr - T.P[q][g][c + 1], T.P[q][g + 2][c + 1] - ctr, S.theta);
                    T.G[3 + q][g][c] = plm_from_differences(ctr - T.P[q][g + 1][c], T.P[q][g + 1][c + 2] - ctr, S.theta);
                }
            }
        }
        if (JUMP)
        {
            __syncthreads();
            if (ring_on)
            {
                #pragma unroll
                for (int q = 0; q < 6; ++q) T.G[q][ring_li + 1][ring_lj + 1] = ring_grad[q];
            }
        }
        __syncthreads();

        // ------------------------------------------------------------------ phases 2 + 3
        const double h = mesh.spacing[b], inv_h = mesh.inv_spacing[b];
        const int li0 = STRIP * warp, lj = lane;
        const double dt_over_h = S.dt * inv_h;
        const double yc = 0.5 * (T.yv[lj] + T.yv[lj + 1]);
        const size_t c0 = (size_t(b) * N + (i0 + li0)) * N + (j0 + lj);
        const double* __restrict__ U0 = mesh.U0;
        const double* __restrict__ BR = mesh.br;

        double sums[NUM_SUMS];
        #pragma unroll
        for (int k = 0; k < NUM_SUMS; ++k) sums[k] = 0.0;
        double amax = 0.0;                              // largest signal speed of the updated cells

        // Update of the cell in strip row r from its four face fluxes (block_update_u, scheme.cpp:568-587).
        auto update_cell = [&] (int r, const double* u, const double* u0, double br, const double* un,
                                const double* FxLo, const double* FxHi, const double* FyLo)
        {
            const int li = li0 + r;
            const size_t c = c0 + size_t(r) * N;
            double hy[3];
            #pragma unroll
            for (int q = 0; q < 3; ++q)
            {
                double up = shfl_down1(FyLo[q]);        // the low-y face of lane + 1 is this cell's high-y face
                hy[q] = lane == 31 ? T.YB[q][li] : up;
            }
            const double x = 0.5 * (T.xv[li] + T.xv[li + 1]);
            double src[3], y1, y2;
            if (QMODE) source_terms_q<true>(model, S, x, yc, u[0], u[1], u[2], T.P[1][li + 2][lj + 2], T.P[2][li + 2][lj + 2], u0[0], u0[1], u0[2], br, src, sums, y1, y2, T.sinks[warp]);
            else source_terms<FAST, true>(model, S, x, yc, u[0], u[1], u[2], u0[0], u0[1], u0[2], br, src, sums, y1, y2, T.sinks[warp]);

            double n0 = u[0] - ((FxHi[0] - FxLo[0]) + (hy[0] - FyLo[0])) * dt_over_h + src[0];
            double n1 = u[1] - ((FxHi[1] - FxLo[1]) + (hy[1] - FyLo[1])) * dt_over_h + src[1];
            double n2 = u[2] - ((FxHi[2] - FxLo[2]) + (hy[2] - FyLo[2])) * dt_over_h + src[2];

            if (n0 < 0.0) report_negative(fail, b, (i0 + li) * N + j0 + lj, n0);

            if (combine)
            {
                const double w = 1.0 - S.rk_b0;
                n0 = un[0] * S.rk_b0 + n0 * w;
                n1 = un[1] * S.rk_b0 + n1 * w;
                n2 = un[2] * S.rk_b0 + n2 * w;
            }
            Uout[c] = n0; Uout[FS + c] = n1; Uout[2 * FS + c] = n2;

            if (compute_dt)
            {
                double mx = n1, my = n2;
                if (QMODE) angmom_to_linear_fast(x, yc, n1, n2, mx, my);
                amax = dmax(amax, max_wavespeed<FAST>(model, S, x, yc, y1, y2, n0, mx, my));
            }
        };
        auto load_cell = [&] (int r, double* u, double* u0, double& br, double* un)
        {
            // volatile: issued here, a whole iteration before their use (ptxas otherwise sinks them below the
            // x-face to save registers, and the update then waits on L2)
            const size_t c = c0 + size_t(r) * N;
            auto ldv = [] (const double* p) { double v; asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; };
            u[0] = ldv(Uin + c); u[1] = ldv(Uin + FS + c); u[2] = ldv(Uin + 2 * FS + c);
            br = 0.0; u0[0] = u0[1] = u0[2] = 0.0; un[0] = un[1] = un[2] = 0.0;
            if (has_buffer) { br = ldv(BR + c); u0[0] = ldv(U0 + c); u0[1] = ldv(U0 + FS + c); u0[2] = ldv(U0 + 2 * FS + c); }
            if (combine) { un[0] = ldv(Un + c); un[1] = ldv(Un + FS + c); un[2] = ldv(Un + 2 * FS + c); }
        };

        double u[3], u0[3], un[3], br;
        load_cell(0, u, u0, br, un);        // in flight during the prologue

        // prologue: tile-boundary faces (high-x row by warp 0, high-y column by half of warp 1), then the
        // faces of strip row 0, whose x-flux is also the high-x flux of the strip below
        // JUMP: faces on a block side whose neighbour is finer, one side per warp (0: high-x, 1: high-y, 2: low-y, 3: low-x)
        if (JUMP)
        {
            const int side = warp == 0 ? 1 : (warp == 1 ? 3 : (warp == 2 ? 2 : 0));
            if ((tiC.z >> (1 + side)) & 1)
            {
                double F[3];
                if (side < 2)
                {
                    jump_corrected_face<0, QMODE>(T, mesh, model, S, Uin, Gphys, b, side, lane, i0, j0, N, h, inv_h, F);
                    double (*dst)[SY] = side == 1 ? T.XB[0] + 4 : T.XLo;     // row 4 of XB[q], or XLo[q], q-stride below
                    const int qs = side == 1 ? 5 : 1;
                    dst[0][lane] = F[0]; dst[qs][lane] = F[1]; dst[2 * qs][lane] = F[2];
                }
                else if (lane < SX)
                {
                    jump_corrected_face<1, QMODE>(T, mesh, model, S, Uin, Gphys, b, side, lane, i0, j0, N, h, inv_h, F);
                    double (*dst)[SX] = side == 3 ? T.YB : T.YLo;
                    dst[0][lane] = F[0]; dst[1][lane] = F[1]; dst[2][lane] = F[2];
                }
            }
        }
        if (warp == 0 && ! finer_hi_x)
        {
            double F[3];
            strip_x_face<QMODE>(T, strip_x_eos<FAST>(T, model, S, SX, lj), inv_h, SX, lj, F, &model);
            T.XB[0][4][lj] = F[0]; T.XB[1][4][lj] = F[1]; T.XB[2][4][lj] = F[2];
        }
        else if (warp == 1 && lane < SX && ! finer_hi_y)
        {
            double F[3];
            strip_y_face<QMODE>(T, strip_y_eos<FAST>(T, model, S, lane, SY), inv_h, lane, SY, F, &model);
            T.YB[0][lane] = F[0]; T.YB[1][lane] = F[1]; T.YB[2][lane] = F[2];
        }
        double FxLo[3], FyLo[3];
        strip_x_face<QMODE>(T, strip_x_eos<FAST>(T, model, S, li0, lj), inv_h, li0, lj, FxLo, &model);
        strip_y_face<QMODE>(T, strip_y_eos<FAST>(T, model, S, li0, lj), inv_h, li0, lj, FyLo, &model);

        if (warp > 0) { T.XB[0][warp][lj] = FxLo[0]; T.XB[1][warp][lj] = FxLo[1]; T.XB[2][warp][lj] = FxLo[2]; }
        __syncthreads();
        if (JUMP && finer_lo_y && lane == 0) { FyLo[0] = T.YLo[0][li0]; FyLo[1] = T.YLo[1][li0]; FyLo[2] = T.YLo[2][li0]; }
        if (JUMP && finer_lo_x && warp == 0) { FxLo[0] = T.XLo[0][lj]; FxLo[1] = T.XLo[1][lj]; FxLo[2] = T.XLo[2][lj]; }

        // steady state: the inputs of a row's update are loaded one iteration ahead, at the end of the loop body
        // (live across the back edge, so they cannot be sunk below the face computations that hide their latency)
        #pragma unroll (JUMP ? 1 : STRIP - 1)       // JUMP: rolled -- that variant is instruction-fetch bound (profiles/), the regular one is not
        for (int r = 1; r < STRIP; ++r)
        {
            double FxNew[3], FyNew[3];
            strip_x_face<QMODE>(T, strip_x_eos<FAST>(T, model, S, li0 + r, lj), inv_h, li0 + r, lj, FxNew, &model);
            strip_y_face<QMODE>(T, strip_y_eos<FAST>(T, model, S, li0 + r, lj), inv_h, li0 + r, lj, FyNew, &model);
            if (JUMP && finer_lo_y && lane == 0) { FyNew[0] = T.YLo[0][li0 + r]; FyNew[1] = T.YLo[1][li0 + r]; FyNew[2] = T.YLo[2][li0 + r]; }
            update_cell(r - 1, u, u0, br, un, FxLo, FxNew, FyLo);
            load_cell(r, u, u0, br, un);
            #pragma unroll
            for (int q = 0; q < 3; ++q) { FxLo[q] = FxNew[q]; FyLo[q] = FyNew[q]; }
        }
        // epilogue: the last row's high-x flux is the first face of the next strip (or the boundary row)
        {
            double FxHi[3];
            FxHi[0] = T.XB[0][warp + 1][lj]; FxHi[1] = T.XB[1][warp + 1][lj]; FxHi[2] = T.XB[2][warp + 1][lj];
            update_cell(STRIP - 1, u, u0, br, un, FxLo, FxHi, FyLo);
        }

        // ------------------------------------------------------------------ fold the CTA's sums
        // per warp, then 4 warps through shared memory
        __syncwarp();
        {
            // Eight sums over 32 lanes by recursive halving: at distances 16, 8, 4 a lane keeps half of its values and
            // trades the other half (7 exchanged doubles instead of 8 x 3), then two butterfly steps finish the one value
            // that is left: 9 shuffled doubles per lane instead of 40.  Lane 4 j ends with the total of value j.
            static_assert(NUM_SUMS - GRV_FX == 8, "the halving below is written for eight values");
            double v4[4], v2[2], v1;
            const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
            #pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                const double lo = sums[GRV_FX + k], hi = sums[GRV_FX + 4 + k];
                v4[k] = (b16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, b16 ? lo : hi, 16);
            }
            #pragma unroll
            for (int k = 0; k < 2; ++k) v2[k] = (b8 ? v4[2 + k] : v4[k]) + __shfl_xor_sync(0xffffffffu, b8 ? v4[k] : v4[2 + k], 8);
            v1 = (b4 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, b4 ? v2[0] : v2[1], 4);
            v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
            v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
            if ((lane & 3) == 0) T.red[warp][GRV_FX + (lane >> 2)] = v1;
        }
        if (lane < GRV_FX) T.red[warp][lane] = T.sinks[warp][lane];
        {
            double m = amax;
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = dmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) T.red[warp][NUM_SUMS] = m;
        }
        __syncthreads();
        if (threadIdx.x <= NUM_SUMS)
        {
            const int k = threadIdx.x;
            double* row = partials + size_t(tiC.w) * ROW;
            double a = T.red[0][k], bq = T.red[1][k], cq = T.red[2][k], d = T.red[3][k];
            // min over cells of h / wavespeed = h / max wavespeed: one division per tile
            row[k] = k == NUM_SUMS ? (compute_dt ? h / dmax(dmax(a, bq), dmax(cq, d)) : 1e300) : ((a + bq) + (cq + d)) * (h * h);
        }
    }
}
