/**
 * stage_tma.cuh -- the fused RK-stage kernel for regular blocks (block size a multiple of 32), persistent, with the
 * tiles staged into shared memory by the TMA engine (cp.async.bulk + mbarrier), double-buffered.
 *
 * Same update as stage_strip (phases P1-P8 + P11 of binary::advance_u, Mara3 src/subprog_binary_scheme.cpp:790-904, plus the
 * CFL estimate of :1107-1126 and the RK combination of subprog_binary.cpp:272-275 in the last stage) and the same thread
 * layout (CTA = 4 warps on a 16 x 32 tile, lane <-> column j, warp <-> strip of 4 rows), but:
 *
 *   * the grid is persistent: gridDim.x = SMs x CTAs per SM, CTA c takes tiles c, c + gridDim.x, ... of the launch's tile
 *     list (interior blocks first, blocks with ghost neighbours last), so a launch pays one ramp-up and one drain instead
 *     of one per wave, and the L1 / instruction cache / mbarriers stay warm;
 *   * while tile k is computed, warp 0 has already issued the bulk copies of tile k + 1: per field and row one
 *     cp.async.bulk of the row's cells in the tile's own block (272 or 288 bytes) and, where the tile touches a block side
 *     in y, one 16-byte copy of the two guard cells from the neighbour block.  The copies land in the other half of
 *     P[2][3][20][36] and complete on that half's mbarrier (expect_tx = 17 280 bytes); no thread holds a register or a
 *     scoreboard slot for them, and the only wait left is mbarrier.try_wait at the top of the next tile, which by then
 *     has had a whole tile time (~10 us) to complete;
 *   * conserved -> primitive happens in place in shared memory; the PLM phase keeps TWICE the un-divided difference (the
 *     1/2 of the central slope folds into the face states' half step and the viscous coefficient);
 *   * the arithmetic is regrouped for the fp64 pipe, which bounds this kernel (profiles/): HLLE by side (hlle_viscous_core),
 *     equation of state from pre-scaled masses and a pre-scaled r^2 table, max(0, .) / min(0, .) and the sign test of the new
 *     density on the integer pipes, gravity totals from two running sums per body instead of three, sink and buffer terms
 *     behind warp-uniform per-tile flags.
 *
 * Shared memory 71.9 KB per CTA -> 3 CTAs (12 warps) per SM with up to 168 registers per thread.
 */
#pragma once

namespace m3b { namespace dev { namespace
{
    constexpr unsigned TMA_TILE_BYTES = 3u * (SX + 4) * (SY + 4) * sizeof(double);     // 17 280

    struct tma_smem_t
    {
        double P[2][3][SX + 4][SY + 4];     // raw conserved rows as the bulk copies deliver them, then primitives (in place); two tiles
        double G[6][SX + 2][SY + 2];        // 2 x un-divided PLM differences d/dx (3), d/dy (3) on tile + 1 halo
        double XB[3][5][SY];                // x-face fluxes at rows 4, 8, 12 (strip starts) and 16 (tile boundary)
        double YB[3][SX];                   // y-face fluxes at the tile's high-y boundary
        double cx[2][SX + 2];               // vertex coordinates of the tile (17 + 33 used), two tiles
        double cy[2][SY + 2];
        // squared-distance tables of the tile's face / centre coordinates to the two bodies (k = 0, 1; the y tables hold the
        // softening rs^2) and to the origin (k = 2; FAST: scaled by the viscous coefficient, see eos_face_fast)
        double x2v[3][SX + 1], x2c[3][SX];
        double y2v[3][SY + 1], y2c[3][SY];
        double xc[SX], dxc[2][SX];          // cell-centre x and its distance to the bodies
        double red[STRIP_THREADS / 32][NUM_SUMS + 1];
        double sinks[STRIP_THREADS / 32][8];
        tile_info_t info[2];
        unsigned long long mbar[2];
        int near_sink;
    };

    __device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

    __device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    }
    __device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
    {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "M3B_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra M3B_DONE;\n"
            "bra M3B_WAIT;\n"
            "M3B_DONE:\n"
            "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
    /** cp.async.bulk global -> shared (SASS: UBLKCP); size and both addresses are multiples of 16 bytes */
    __device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
    {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
    }

    /**
     * Warp 0: start the bulk copies of tile `tile` into half `buf` and fetch the tile's vertex coordinates into registers
     * (nx: xv[i0 + lane] for lane <= 16, ny: yv[j0 + lane], ny32: yv[j0 + 32] on lane 0; stored to shared memory later,
     * when they have arrived).  60 (field, row) pairs over 32 lanes; a row is one copy from the tile's own block column
     * range plus 16-byte copies of the two guard columns beyond a block side.
     */
    __device__ __forceinline__ void tma_issue_tile(tma_smem_t& T, const mesh_dev_t& mesh, const tile_info_t* __restrict__ tile_info,
        const double* __restrict__ Uin, int tile, int buf, int N, int tiles_y, int tpb, int lane, double& nx, double& ny, double& ny32)
    {
        // multi-GPU: tiles of blocks with ghost neighbours may only be fetched once the guard-zone unpack has finished
        if (tile >= mesh.first_wait_cta)
        {
            if (lane == 0)
            {
                unsigned long long v;
                do { asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mesh.ready_flag) : "memory"); } while (v < mesh.ready_value);
            }
            __syncwarp();
        }
        const int4* ti4 = reinterpret_cast<const int4*>(tile_info + tile);
        const int4 tiA = __ldg(ti4), tiB = __ldg(ti4 + 1), tiC = __ldg(ti4 + 2);
        if (lane == 0)
        {
            int4* d = reinterpret_cast<int4*>(&T.info[buf]);
            d[0] = tiA; d[1] = tiB; d[2] = tiC;
            mbar_arrive_expect_tx(&T.mbar[buf], TMA_TILE_BYTES);
        }
        __syncwarp();
        const int b = tiA.x, t = tile % tpb;
        const int i0 = (t / tiles_y) * SX, j0 = (t % tiles_y) * SY;
        const size_t FS = mesh.FS;
        {
            const double* xvg = mesh.xv + size_t(b) * (N + 1) + i0;
            const double* yvg = mesh.yv + size_t(b) * (N + 1) + j0;
            nx = lane <= SX ? __ldg(xvg + lane) : 0.0;
            ny = __ldg(yvg + lane);
            ny32 = lane == 0 ? __ldg(yvg + SY) : 0.0;
        }
        const int* n9 = T.info[buf].n9;
        const int c0 = j0 == 0 ? 0 : j0 - 2, c1 = j0 + SY == N ? N : j0 + SY + 2;      // the row's columns inside the tile's own block column
        const int dcol = c0 - (j0 - 2);
        const unsigned main_bytes = unsigned(c1 - c0) * sizeof(double);
        #pragma unroll
        for (int idx = lane; idx < 3 * (SX + 4); idx += 32)
        {
            const int f = idx / (SX + 4), row = idx - f * (SX + 4);
            const int gi = i0 - 2 + row;
            const int di = gi < 0 ? -1 : (gi >= N ? 1 : 0);
            const long ii = gi - di * N;
            const double* field = Uin + f * FS;
            double* drow = &T.P[buf][f][row][0];
            bulk_copy_g2s(drow + dcol, field + (long(n9[(di + 1) * 3 + 1]) * N + ii) * N + c0, main_bytes, &T.mbar[buf]);
            if (j0 == 0)      bulk_copy_g2s(drow,          field + (long(n9[(di + 1) * 3 + 0]) * N + ii) * N + (N - 2), 16u, &T.mbar[buf]);
            if (j0 + SY == N) bulk_copy_g2s(drow + SY + 2, field + (long(n9[(di + 1) * 3 + 2]) * N + ii) * N,           16u, &T.mbar[buf]);
        }
    }

    /** x-face between tile cells (li - 1, lj) and (li, lj), 0 <= li <= SX; yd[k] = T.y2c[k][lj] */
    template<bool FAST>
    __device__ __forceinline__ void tma_x_face(const tma_smem_t& T, const double (*P)[SX + 4][SY + 4], const model_t& model, const stage_t& S,
        const strip_consts_t& C, double cvis, const double yd[3], int li, int lj, double F[3])
    {
        const double d1 = T.x2v[0][li] + yd[0], d2 = T.x2v[1][li] + yd[1], q2 = T.x2v[2][li] + yd[2];
        eos_face_t e;
        if (FAST) e = eos_face_fast(C, d1, d2, q2);
        else { const eos_t g = eos_from_distances<false>(model, S, d1, d2, q2); e.cs2 = g.cs2; e.cs = g.cs; e.mu_coef = cvis * g.nu; }
        const prim_t L = {fma(T.G[0][li][lj + 1], 0.25, P[0][li + 1][lj + 2]), fma(T.G[1][li][lj + 1], 0.25, P[1][li + 1][lj + 2]), fma(T.G[2][li][lj + 1], 0.25, P[2][li + 1][lj + 2])};
        const prim_t R = {fma(T.G[0][li + 1][lj + 1], -0.25, P[0][li + 2][lj + 2]), fma(T.G[1][li + 1][lj + 1], -0.25, P[1][li + 2][lj + 2]), fma(T.G[2][li + 1][lj + 1], -0.25, P[2][li + 2][lj + 2])};
        hlle_viscous_core<0>(e.cs2, e.cs, e.mu_coef, L, R,
            T.G[1][li][lj + 1] + T.G[1][li + 1][lj + 1], T.G[2][li][lj + 1] + T.G[2][li + 1][lj + 1],
            T.G[4][li][lj + 1] + T.G[4][li + 1][lj + 1], T.G[5][li][lj + 1] + T.G[5][li + 1][lj + 1], F);
    }

    /** y-face between tile cells (li, lj - 1) and (li, lj), 0 <= lj <= SY; yd[k] = T.y2v[k][lj] */
    template<bool FAST>
    __device__ __forceinline__ void tma_y_face(const tma_smem_t& T, const double (*P)[SX + 4][SY + 4], const model_t& model, const stage_t& S,
        const strip_consts_t& C, double cvis, const double yd[3], int li, int lj, double F[3])
    {
        const double d1 = T.x2c[0][li] + yd[0], d2 = T.x2c[1][li] + yd[1], q2 = T.x2c[2][li] + yd[2];
        eos_face_t e;
        if (FAST) e = eos_face_fast(C, d1, d2, q2);
        else { const eos_t g = eos_from_distances<false>(model, S, d1, d2, q2); e.cs2 = g.cs2; e.cs = g.cs; e.mu_coef = cvis * g.nu; }
        const prim_t L = {fma(T.G[3][li + 1][lj], 0.25, P[0][li + 2][lj + 1]), fma(T.G[4][li + 1][lj], 0.25, P[1][li + 2][lj + 1]), fma(T.G[5][li + 1][lj], 0.25, P[2][li + 2][lj + 1])};
        const prim_t R = {fma(T.G[3][li + 1][lj + 1], -0.25, P[0][li + 2][lj + 2]), fma(T.G[4][li + 1][lj + 1], -0.25, P[1][li + 2][lj + 2]), fma(T.G[5][li + 1][lj + 1], -0.25, P[2][li + 2][lj + 2])};
        hlle_viscous_core<1>(e.cs2, e.cs, e.mu_coef, L, R,
            T.G[4][li + 1][lj] + T.G[4][li + 1][lj + 1], T.G[5][li + 1][lj] + T.G[5][li + 1][lj + 1],
            T.G[1][li + 1][lj] + T.G[1][li + 1][lj + 1], T.G[2][li + 1][lj] + T.G[2][li + 1][lj + 1], F);
    }

    template<int MIN_CTAS, int NB, bool FAST, int MODE>
    __global__ void __launch_bounds__(STRIP_THREADS, MIN_CTAS) stage_tma(
        mesh_dev_t mesh, model_t model, const stage_t* __restrict__ stage_ptr, const tile_info_t* __restrict__ tile_info, int num_tiles,
        const double* __restrict__ Uin, const double* __restrict__ Un, double* __restrict__ Uout, double* partials, fail_dev_t* fail)
    {
        extern __shared__ __align__(128) unsigned char smem_raw[];
        tma_smem_t& T = *reinterpret_cast<tma_smem_t*>(smem_raw);

        const stage_t S = *stage_ptr;       // written by the host or by prepare_next of the step before
        // MODE 1 / 2: first / last stage of an RK2 step with adaptive dt, flags known at compile time; 0: read them from S
        const bool combine = MODE == 0 ? S.combine != 0 : MODE == 2, compute_dt = MODE == 0 ? S.compute_dt != 0 : MODE == 2;
        const int N = NB ? NB : mesh.N;
        const int tiles_y = N / SY, tpb = (N / SX) * tiles_y;
        const size_t FS = mesh.FS;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const strip_consts_t C = {S.m1 * model.inv_mach2, S.m2 * model.inv_mach2, -S.m1, -S.m2, 2.0 * S.theta, S.dt};
        const double* __restrict__ U0 = mesh.U0;
        const double* __restrict__ BR = mesh.br;

        if (threadIdx.x == 0)
        {
            mbar_init(&T.mbar[0], 1);
            mbar_init(&T.mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();

        int tile = blockIdx.x;
        double nx = 0.0, ny = 0.0, ny32 = 0.0;
        if (warp == 0 && tile < num_tiles)
        {
            tma_issue_tile(T, mesh, tile_info, Uin, tile, 0, N, tiles_y, tpb, lane, nx, ny, ny32);
            if (lane <= SX) T.cx[0][lane] = nx;
            T.cy[0][lane] = ny;
            if (lane == 0) T.cy[0][SY] = ny32;
        }
        __syncthreads();

        for (int k = 0; tile < num_tiles; ++k, tile += gridDim.x)
        {
            const int buf = k & 1;
            const int next = tile + int(gridDim.x);
            const bool has_next = next < num_tiles;
            // the other half was released by the barrier that ended tile k - 1: refill it while this tile is computed
            if (warp == 0 && has_next)
            {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tma_issue_tile(T, mesh, tile_info, Uin, next, buf ^ 1, N, tiles_y, tpb, lane, nx, ny, ny32);
            }

            const int b = T.info[buf].b, flags = T.info[buf].flags;
            const int t = tile % tpb;
            const int i0 = (t / tiles_y) * SX, j0 = (t % tiles_y) * SY;
            const bool has_buffer = flags & 1;
            const double h = mesh.spacing[b], inv_h = mesh.inv_spacing[b];
            const double cvis = 0.125 * inv_h;      // 0.5 nu x face average 0.5 x (1 / 2h) of the doubled differences

            // the update phase's inputs are first touched several us from now: pull their lines into L2 already
            if (lane < 2 * STRIP)
            {
                const size_t c = (size_t(b) * N + (i0 + STRIP * warp + (lane >> 1))) * N + j0 + 16 * (lane & 1);
                if (has_buffer)
                {
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(BR + c));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(U0 + c));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(U0 + FS + c));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(U0 + 2 * FS + c));
                }
                if (combine)
                {
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(Un + c));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(Un + FS + c));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(Un + 2 * FS + c));
                }
            }
            if (lane < 8) T.sinks[warp][lane] = 0.0;

            // coordinate tables (warps 1-3; warp 0 has just issued the next tile)
            {
                const double bx[3] = {S.x1, S.x2, 0.0}, by[3] = {S.y1, S.y2, 0.0};
                const double soft[3] = {model.softening_radius2, model.softening_radius2, 0.0};
                // FAST: the origin table carries the viscous coefficient c = cvis alpha / Mach: (c x)^2 + (c y)^2 = (c r)^2
                const double cf = FAST ? cvis * model.alpha * model.inv_mach : 1.0;
                if (warp == 1)
                {
                    if (lane <= SX)
                    {
                        const double xv = T.cx[buf][lane];
                        #pragma unroll
                        for (int q = 0; q < 3; ++q) { const double d = (xv - bx[q]) * (q == 2 ? cf : 1.0); T.x2v[q][lane] = d * d; }
                    }
                    if (lane < SX)
                    {
                        const double xc = 0.5 * (T.cx[buf][lane] + T.cx[buf][lane + 1]);
                        T.xc[lane] = xc;
                        T.dxc[0][lane] = xc - bx[0];
                        T.dxc[1][lane] = xc - bx[1];
                        #pragma unroll
                        for (int q = 0; q < 3; ++q) { const double d = (xc - bx[q]) * (q == 2 ? cf : 1.0); T.x2c[q][lane] = d * d; }
                    }
                }
                else if (warp == 2)
                {
                    const double yv = T.cy[buf][lane], yc = 0.5 * (yv + T.cy[buf][lane + 1]);
                    #pragma unroll
                    for (int q = 0; q < 3; ++q)
                    {
                        const double dv = (yv - by[q]) * (q == 2 ? cf : 1.0), dc = (yc - by[q]) * (q == 2 ? cf : 1.0);
                        T.y2v[q][lane] = fma(dv, dv, soft[q]);
                        T.y2c[q][lane] = fma(dc, dc, soft[q]);
                    }
                }
                else if (warp == 3 && lane == 0)
                {
                    const double yv = T.cy[buf][SY];
                    #pragma unroll
                    for (int q = 0; q < 3; ++q) { const double dv = (yv - by[q]) * (q == 2 ? cf : 1.0); T.y2v[q][SY] = fma(dv, dv, soft[q]); }
                    // does any cell of the tile lie within the sinks' reach (a2 = dr^2 / (2 s^2) < 100)?  distance of each body to the tile's rectangle
                    const double xlo = T.cx[buf][0], xhi = T.cx[buf][SX], ylo = T.cy[buf][0], yhi = yv;
                    bool near = false;
                    #pragma unroll
                    for (int q = 0; q < 2; ++q)
                    {
                        const double ddx = dmax(dmax(xlo - bx[q], bx[q] - xhi), 0.0), ddy = dmax(dmax(ylo - by[q], by[q] - yhi), 0.0);
                        near = near || (ddx * ddx + ddy * ddy) * model.sink_inv_2s2 < 100.0;
                    }
                    T.near_sink = near;
                }
            }

            // ------------------------------------------------------------------ phase 0: wait for the tile, primitives in place
            mbar_wait(&T.mbar[buf], (k >> 1) & 1);
            double (*P)[SX + 4][SY + 4] = T.P[buf];
            {
                // 20 rows of 18 sixteen-byte chunks (two cells in y): thread <-> (row rr + 7 m, chunk cc), m = 0, 1, 2
                const int rr = threadIdx.x / 18, cc = threadIdx.x - 18 * rr;
                if (rr < 7)
                {
                    #pragma unroll
                    for (int m = 0; m < 3; ++m)
                    {
                        const int row = rr + 7 * m;
                        if (row < SX + 4)
                        {
                            // iso2d::recover_primitive (physics_iso2d.hpp:351-362) for the chunk's two cells
                            const double2 u0 = *reinterpret_cast<const double2*>(&P[0][row][2 * cc]);
                            const double2 u1 = *reinterpret_cast<const double2*>(&P[1][row][2 * cc]);
                            const double2 u2 = *reinterpret_cast<const double2*>(&P[2][row][2 * cc]);
                            const double ia = fast_rcp(u0.x), ib = fast_rcp(u0.y);
                            *reinterpret_cast<double2*>(&P[1][row][2 * cc]) = make_double2(u1.x * ia, u1.y * ib);
                            *reinterpret_cast<double2*>(&P[2][row][2 * cc]) = make_double2(u2.x * ia, u2.y * ib);
                        }
                    }
                }
            }
            __syncthreads();

            // ------------------------------------------------------------------ phase 1: PLM differences (doubled)
            {
                // gradient rows g = 0..17 (<-> P row g + 1): warps take 5, 5, 4, 4 consecutive rows; lane <-> gradient column
                // c = lane (<-> P column lane + 1).  Marching down the rows the backward x-difference of a row is the forward
                // difference of the row before.
                const int g0 = warp < 2 ? 5 * warp : 10 + 4 * (warp - 2);
                const int g1 = g0 + (warp < 2 ? 5 : 4);
                double pc[3], dl[3];
                #pragma unroll
                for (int q = 0; q < 3; ++q) { pc[q] = P[q][g0 + 1][lane + 1]; dl[q] = pc[q] - P[q][g0][lane + 1]; }

                #pragma unroll 2
                for (int g = g0; g < g1; ++g)
                {
                    double pp[3], yl[3], yr[3];
                    #pragma unroll
                    for (int q = 0; q < 3; ++q) { pp[q] = P[q][g + 2][lane + 1]; yl[q] = P[q][g + 1][lane]; yr[q] = P[q][g + 1][lane + 2]; }
                    #pragma unroll
                    for (int q = 0; q < 3; ++q)
                    {
                        const double dr = pp[q] - pc[q];
                        T.G[q][g][lane]     = plm2_from_differences(dl[q], dr, C.theta2);
                        T.G[3 + q][g][lane] = plm2_from_differences(pc[q] - yl[q], yr[q] - pc[q], C.theta2);
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
                        const double ctr = P[q][g + 1][c + 1];
                        T.G[q][g][c]     = plm2_from_differences(ctr - P[q][g][c + 1], P[q][g + 2][c + 1] - ctr, C.theta2);
                        T.G[3 + q][g][c] = plm2_from_differences(ctr - P[q][g + 1][c], P[q][g + 1][c + 2] - ctr, C.theta2);
                    }
                }
                // the next tile's vertex coordinates have arrived long ago
                if (warp == 0 && has_next)
                {
                    if (lane <= SX) T.cx[buf ^ 1][lane] = nx;
                    T.cy[buf ^ 1][lane] = ny;
                    if (lane == 0) T.cy[buf ^ 1][SY] = ny32;
                }
            }
            __syncthreads();

            // ------------------------------------------------------------------ phases 2 + 3: faces, update
            const int li0 = STRIP * warp, lj = lane;
            const double dt_over_h = S.dt * inv_h;
            const double yc = 0.5 * (T.cy[buf][lj] + T.cy[buf][lj + 1]);
            const double dy1 = yc - S.y1, dy2 = yc - S.y2;
            const double ydc[3] = {T.y2c[0][lj], T.y2c[1][lj], T.y2c[2][lj]};      // x-faces and cell centres share the column's y-part
            const double ydv[3] = {T.y2v[0][lj], T.y2v[1][lj], T.y2v[2][lj]};
#ifdef M3B_HOT_PATH_ONLY        // tools/hot_path_count.sh: the instruction census of the common path (no sink within reach, no negative density)
            const bool near_sink = false;
#else
            const bool near_sink = T.near_sink != 0;
#endif
            const size_t c0 = (size_t(b) * N + (i0 + li0)) * N + (j0 + lj);

            strip_sums_t sums = {{0.0, 0.0}, {0.0, 0.0}, 0.0, 0.0, 0.0};
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
                const double x = T.xc[li];
                double acc[3], y1, y2;
                source_terms_strip<FAST>(model, C, x, yc, T.dxc[0][li], dy1, T.dxc[1][li], dy2, T.x2c[0][li] + ydc[0], T.x2c[1][li] + ydc[1],
                    near_sink, has_buffer, u[0], u[1], u[2], u0[0], u0[1], u0[2], br, acc, sums, y1, y2, T.sinks[warp]);

                double n0 = fma(-((FxHi[0] - FxLo[0]) + (hy[0] - FyLo[0])), dt_over_h, acc[0]);
                double n1 = fma(-((FxHi[1] - FxLo[1]) + (hy[1] - FyLo[1])), dt_over_h, acc[1]);
                double n2 = fma(-((FxHi[2] - FxLo[2]) + (hy[2] - FyLo[2])), dt_over_h, acc[2]);

#ifndef M3B_HOT_PATH_ONLY
                if (__double2hiint(n0) < 0) { if (n0 < 0.0) report_negative(fail, b, (i0 + li) * N + j0 + lj, n0); }
#endif

                if (combine)
                {
                    const double w = 1.0 - S.rk_b0;
                    n0 = un[0] * S.rk_b0 + n0 * w;
                    n1 = un[1] * S.rk_b0 + n1 * w;
                    n2 = un[2] * S.rk_b0 + n2 * w;
                }
                Uout[c] = n0; Uout[FS + c] = n1; Uout[2 * FS + c] = n2;

                if (compute_dt)
                    amax = dmax(amax, FAST ? max_wavespeed_fast(C, y1, y2, n0, n1, n2) : max_wavespeed<false>(model, S, x, yc, y1, y2, n0, n1, n2));
            };
            auto load_cell = [&] (int r, double* u, double* u0, double& br, double* un)
            {
                // volatile: issued here, a whole iteration before their use
                const size_t c = c0 + size_t(r) * N;
                auto ldv = [] (const double* p) { double v; asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; };
                u[0] = ldv(Uin + c); u[1] = ldv(Uin + FS + c); u[2] = ldv(Uin + 2 * FS + c);
                br = 0.0; u0[0] = u0[1] = u0[2] = 0.0; un[0] = un[1] = un[2] = 0.0;
                if (has_buffer) { br = ldv(BR + c); u0[0] = ldv(U0 + c); u0[1] = ldv(U0 + FS + c); u0[2] = ldv(U0 + 2 * FS + c); }
                if (combine) { un[0] = ldv(Un + c); un[1] = ldv(Un + FS + c); un[2] = ldv(Un + 2 * FS + c); }
            };

            double u[3], u0[3], un[3], br;
            load_cell(0, u, u0, br, un);        // in flight during the prologue

            // prologue: tile-boundary faces (high-x row by warp 0, high-y column by half of warp 1), then the faces of strip
            // row 0, whose x-flux is also the high-x flux of the strip below
            if (warp == 0)
            {
                double F[3];
                tma_x_face<FAST>(T, P, model, S, C, cvis, ydc, SX, lj, F);
                T.XB[0][4][lj] = F[0]; T.XB[1][4][lj] = F[1]; T.XB[2][4][lj] = F[2];
            }
            else if (warp == 1 && lane < SX)
            {
                double F[3];
                const double ydhi[3] = {T.y2v[0][SY], T.y2v[1][SY], T.y2v[2][SY]};
                tma_y_face<FAST>(T, P, model, S, C, cvis, ydhi, lane, SY, F);
                T.YB[0][lane] = F[0]; T.YB[1][lane] = F[1]; T.YB[2][lane] = F[2];
            }
            double FxLo[3], FyLo[3];
            tma_x_face<FAST>(T, P, model, S, C, cvis, ydc, li0, lj, FxLo);
            tma_y_face<FAST>(T, P, model, S, C, cvis, ydv, li0, lj, FyLo);

            if (warp > 0) { T.XB[0][warp][lj] = FxLo[0]; T.XB[1][warp][lj] = FxLo[1]; T.XB[2][warp][lj] = FxLo[2]; }
            __syncthreads();

            #pragma unroll
            for (int r = 1; r < STRIP; ++r)
            {
                double FxNew[3], FyNew[3];
                tma_x_face<FAST>(T, P, model, S, C, cvis, ydc, li0 + r, lj, FxNew);
                tma_y_face<FAST>(T, P, model, S, C, cvis, ydv, li0 + r, lj, FyNew);
                update_cell(r - 1, u, u0, br, un, FxLo, FxNew, FyLo);
                load_cell(r, u, u0, br, un);
                #pragma unroll
                for (int q = 0; q < 3; ++q) { FxLo[q] = FxNew[q]; FyLo[q] = FyNew[q]; }
            }
            {
                double FxHi[3];
                FxHi[0] = T.XB[0][warp + 1][lj]; FxHi[1] = T.XB[1][warp + 1][lj]; FxHi[2] = T.XB[2][warp + 1][lj];
                update_cell(STRIP - 1, u, u0, br, un, FxLo, FxHi, FyLo);
            }

            // ------------------------------------------------------------------ fold the CTA's sums
            __syncwarp();
            {
                // the six gravity totals and the ejected angular momentum from the thread's running sums (strip_sums_t)
                double v8[8];
                v8[0] = fma(-S.x1, sums.S0[0], sums.Sx[0]);               v8[1] = fma(-S.x2, sums.S0[1], sums.Sx[1]);                 // GRV_FX
                v8[2] = dy1 * sums.S0[0];                                 v8[3] = dy2 * sums.S0[1];                                   // GRV_FY
                v8[4] = fma(S.x1 * yc, sums.S0[0], -S.y1 * sums.Sx[0]);   v8[5] = fma(S.x2 * yc, sums.S0[1], -S.y2 * sums.Sx[1]);     // GRV_TQ
                v8[6] = sums.buf_m;                                       v8[7] = fma(-yc, sums.buf_px, sums.buf_xpy);                // BUF_M, BUF_L
                static_assert(GRV_FX == 8 && GRV_FY == 10 && GRV_TQ == 12 && BUF_M == 14 && BUF_L == 15 && NUM_SUMS == 16, "layout of the eight values");
                // Eight sums over 32 lanes by recursive halving: at distances 16, 8, 4 a lane keeps half of its values and trades
                // the other half, then two butterfly steps finish the one value that is left.  Lane 4 j ends with value j.
                double v4[4], v2[2], v1;
                const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
                #pragma unroll
                for (int q = 0; q < 4; ++q)
                {
                    const double lo = v8[q], hi = v8[4 + q];
                    v4[q] = (b16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, b16 ? lo : hi, 16);
                }
                #pragma unroll
                for (int q = 0; q < 2; ++q) v2[q] = (b8 ? v4[2 + q] : v4[q]) + __shfl_xor_sync(0xffffffffu, b8 ? v4[q] : v4[2 + q], 8);
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
            __syncthreads();        // ends the tile: P[buf], G, XB / YB and the tables are free again
            if (threadIdx.x <= NUM_SUMS)
            {
                const int q = threadIdx.x;
                double* row = partials + size_t(tile) * ROW;
                const double a = T.red[0][q], bq = T.red[1][q], cq = T.red[2][q], d = T.red[3][q];
                // min over cells of h / wavespeed = h / max wavespeed: one division per tile
                row[q] = q == NUM_SUMS ? (compute_dt ? h / dmax(dmax(a, bq), dmax(cq, d)) : 1e300) : ((a + bq) + (cq + d)) * (h * h);
            }
        }
    }
}}}
