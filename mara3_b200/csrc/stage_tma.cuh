/**
 * stage_tma.cuh -- the fused RK-stage kernel for regular blocks (block size a multiple of 32): persistent CTAs, tiles staged
 * into shared memory asynchronously (cp.async, SASS LDGSTS) one tile ahead of the arithmetic, double-buffered.
 *
 * Same update as stage_strip (phases P1-P8 + P11 of binary::advance_u, Mara3 src/subprog_binary_scheme.cpp:790-904, plus the
 * CFL estimate of :1107-1126 and the RK combination of subprog_binary.cpp:272-275 in the last stage) and the same thread
 * layout (CTA = 4 warps on a 16 x 32 tile, lane <-> column j, warp <-> strip of 4 rows), but:
 *
 *   * the grid is persistent: gridDim.x = SMs x CTAs per SM, CTA c takes tiles c, c + gridDim.x, ... of the launch's tile
 *     list (interior blocks first, blocks with ghost neighbours last), so a launch pays one ramp-up and one drain instead
 *     of one per wave;
 *   * while tile k is computed, the 16-byte chunks of tile k + 1 (20 rows x 18 chunks x 3 fields, each thread the <= 9 chunks
 *     it will convert itself) are already in flight into the other half of P[2][3][20][36]: no register and no scoreboard
 *     slot is held for them, and a thread only waits (cp.async.wait_group) for its OWN chunks at the top of the next tile,
 *     a whole tile time (~10 us) after it asked for them;
 *   * conserved -> primitive happens in place in shared memory; the PLM phase keeps TWICE the un-divided difference (the
 *     1/2 of the central slope folds into the face states' half step and the viscous coefficient);
 *   * the arithmetic is regrouped for the fp64 pipe, which bounds this kernel (profiles/): HLLE by side (hlle_viscous_core),
 *     equation of state from pre-scaled masses and a pre-scaled r^2 table, max(0, .) / min(0, .) and the sign test of the new
 *     density on the integer pipes, gravity totals from two running sums per body instead of three, sink and buffer terms
 *     behind warp-uniform per-tile flags.
 *
 * A first version moved the rows with cp.async.bulk (UBLKCP, one copy per field and row): UBLKCP takes uniform-register
 * operands, so per-lane addresses turn into a waterfall loop of ~8 instructions per copy on the issuing warp, ~120 copies per
 * tile; the other warps waited for it at the barriers (793 us against stage_strip's 565 us on 4096^2,
 * profiles/r02_stage_tma_bulk_rows_ncu_summary.txt).  Rows of 36 doubles that change block at every side are the wrong
 * granularity for the TMA engine; 16-byte cp.async per thread is the right one.
 *
 * Shared memory 72 KB per CTA -> 3 CTAs (12 warps) per SM with up to 168 registers per thread.
 */
#pragma once

namespace m3b { namespace dev { namespace
{
// the update inputs' load instructions (M3B_UPDATE_LD_NC: bit 0 the cell's own state, bit 1 initial state and buffer rate, bit 2 the
// step-start state): owned cells of arrays that nothing writes during the launch, so ld.global.nc is allowed -- and lets ptxas
// lift them over the stores of the rows before
// Measured on 4096^2 (us per launch): none 571, own state 554, initial state + rate 557, both 563, step-start state 562, all 562.
#ifndef M3B_UPDATE_LD_NC
#define M3B_UPDATE_LD_NC 1
#endif
#define M3B_LD_SEL(bit) (((M3B_UPDATE_LD_NC) >> (bit)) & 1)
#ifndef M3B_PLM_UNROLL
#define M3B_PLM_UNROLL 2
#endif
    constexpr int PLM_UNROLL = M3B_PLM_UNROLL;      // rows of the PLM phase per loop body

    template<int NBUF>
    struct tma_smem_t
    {
        double P[NBUF][3][SX + 4][SY + 4];     // raw conserved rows as the bulk copies deliver them, then primitives (in place); two tiles
        double G[6][SX + 2][SY + 2];        // 2 x un-divided PLM differences d/dx (3), d/dy (3) on tile + 1 halo
        double XB[3][4][SY];                // x-face fluxes at rows 4, 8, 12 (strip starts) and 16 (tile boundary): the high-x flux of strip w is XB[.][w]
        double YB[3][SX];                   // y-face fluxes at the tile's high-y boundary
        double cx[2][SX + 2];               // vertex coordinates of the tile (17 + 33 used), two tiles
        double cy[2][SY + 2];
        double hh[2][2];                    // spacing and 1 / spacing of the tile's block, two tiles
        // squared-distance tables of the tile's face / centre coordinates to the two bodies (k = 0, 1; the y tables hold the
        // softening rs^2) and to the origin (k = 2; FAST: scaled by the viscous coefficient, see eos_face_fast)
        double x2v[3][SX + 1], x2c[3][SX];
        double y2v[3][SY + 1], y2c[3][SY];
        double xc[SX], dxc[2][SX];          // cell-centre x and its distance to the bodies
        double red[STRIP_THREADS / 32][NUM_SUMS + 1];
        double sinks[STRIP_THREADS / 32][8];
        tile_info_t info[3];                // records of the tiles k, k + 1, k + 2 (slot = ordinal % 3)
        int near_sink;
        int unpack_slot;                    // exchange_unpack: the strip this CTA has taken off the list
    };

    /** The warp's role in the tile (strip of rows, boundary faces, tables), shifted by the third of the grid the CTA lies in: warp w
     *  of every CTA sits on scheduler w, warps 0 and 1 carry the tile-boundary faces, and the three CTAs of an SM come one from each
     *  third of a full grid -- unshifted, one scheduler would hold the three heaviest warps of the SM (measured on 4096^2: 566 ->
     *  562 us per launch; -DM3B_NO_ROTATE_ROLES for the comparison). */
    __device__ __forceinline__ int logical_warp()
    {
#ifdef M3B_NO_ROTATE_ROLES
        return threadIdx.x >> 5;
#else
        return int((threadIdx.x >> 5) + blockIdx.x * 3u / gridDim.x) & 3;
#endif
    }

    __device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

    /** cp.async 16 bytes global -> shared, L2 only (SASS: LDGSTS.E.BYPASS.128) */
    __device__ __forceinline__ void cp_async_16(void* dst, const void* src)
    {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
    }
    __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    __device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

    /**
     * Every thread: ask for the chunks of the tile described by `ti` that this thread will convert -- (row rr + 7 m, chunk cc),
     * m = 0, 1, 2, of the (SX + 4) x (SY + 4) region = 20 rows of 18 sixteen-byte chunks (two cells in y; columns j0 - 2 and N
     * are even, so a chunk never straddles two blocks), all three fields -- into half `buf`.
     */
    template<bool PREFETCH_ONLY = false, typename SMEM>
    __device__ __forceinline__ void stage_tile_async(SMEM& T, const tile_info_t& ti, const double* __restrict__ Uin, size_t FS,
        int t, int buf, int N, int tiles_y)
    {
        const int rr = threadIdx.x / 18, cc = threadIdx.x - 18 * rr;
        if (rr < 7)
        {
            const int i0 = (t / tiles_y) * SX, j0 = (t % tiles_y) * SY;
            const int gj = j0 - 2 + 2 * cc;
            const int dj = gj < 0 ? -1 : (gj >= N ? 1 : 0);
            const long col = gj - dj * N;
            const int nlo = ti.n9[dj + 1], nmid = ti.n9[3 + dj + 1], nhi = ti.n9[6 + dj + 1];
            #pragma unroll
            for (int m = 0; m < 3; ++m)
            {
                const int row = rr + 7 * m;
                if (row < SX + 4)
                {
                    const int gi = i0 - 2 + row;
                    const int di = gi < 0 ? -1 : (gi >= N ? 1 : 0);
                    const int nb = di < 0 ? nlo : (di > 0 ? nhi : nmid);
                    const double* src = Uin + (long(nb) * N + (gi - di * N)) * N + col;
                    if (PREFETCH_ONLY)
                    {
                        // one buffer only (4 CTAs per SM): the tile cannot land yet, but its lines can already come from HBM into L2
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(src));
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(src + FS));
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(src + 2 * FS));
                    }
                    else
                    {
                        cp_async_16(&T.P[buf][0][row][2 * cc], src);
                        cp_async_16(&T.P[buf][1][row][2 * cc], src + FS);
                        cp_async_16(&T.P[buf][2][row][2 * cc], src + 2 * FS);
                    }
                }
            }
        }
        if (! PREFETCH_ONLY) cp_async_commit();
    }

    /**
     * Send side of extend() across GPUs (scheme.cpp:132-142), first thing in the kernel: the CTAs copy the strips / corners of
     * owned blocks straight into the destination ranks' landing buffers (entry.pad = destination rank, entry.offset = position
     * in ITS buffer); the CTA that completes the last strip raises this rank's flag on every destination.
     */
    __device__ __forceinline__ void exchange_push(const fused_exchange_t& X, size_t FS, int N, int* slot)
    {
        int* cnt = X.counters + 4 * X.cset;
        if (blockIdx.x == 0 && threadIdx.x < 4) X.counters[4 * (1 - X.cset) + threadIdx.x] = 0;       // the set the previous fused launch used
        unsigned long long t0 = 0;
        if (threadIdx.x == 0 && X.clock_words) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;)
        {
            // strips are taken off the list, not dealt out: the flag below must not wait for a CTA that is not resident yet
            if (threadIdx.x == 0) *slot = atomicAdd(cnt + 3, 1);
            __syncthreads();
            const int n = *slot;
            __syncthreads();
            if (n >= X.n_push) break;
            const halo_entry_dev_t e = X.push[n];
            const int cells = e.ni * e.nj;
            double* __restrict__ dst = X.peers.recv[e.pad][X.parity] + e.offset;
            for (int k = threadIdx.x; k < 3 * cells; k += blockDim.x)
            {
                const int q = k / cells, c = k % cells;
                const int i = e.i0 + c / e.nj, j = e.j0 + c % e.nj;
                dst[k] = X.U[q * FS + (size_t(e.block) * N + i) * N + j];
            }
            __threadfence_system();
            __syncthreads();
            if (threadIdx.x == 0 && atomicAdd(cnt, 1) == X.n_push - 1)
            {
                // that was the last strip: everything this rank sends has left
                __threadfence_system();
                for (int p = 0; p < MAX_PEERS; ++p)
                    if ((X.dest_mask >> p) & 1u) store_release_sys(X.peers.halo_flag[p] + X.me, X.counter);
                if (X.clock_words)
                {
                    // kernel start (this CTA's) -> every strip stored in the neighbours' memory and the flags raised
                    unsigned long long t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    atomicAdd(X.clock_words + 4, t1 - t0);
                    atomicAdd(X.clock_words + 5, 1ull);
                }
            }
        }
    }

    /**
     * Receive side, called by a whole CTA before it fetches its first tile of a block with ghost neighbours: take strips off
     * the list (atomic counter) while there are any -- wait for the source rank's flag, scatter the landed strip into the ghost
     * block -- then wait until every strip is in place.  All CTAs of the grid are resident (persistent launch), the neighbours'
     * kernels push before they do anything else, and nothing here waits for a kernel that is not running: no deadlock.
     */
    __device__ __forceinline__ void exchange_unpack(const fused_exchange_t& X, size_t FS, int N, int* slot)
    {
        int* cnt = X.counters + 4 * X.cset;
        unsigned long long t0 = 0;
        if (threadIdx.x == 0 && X.clock_words) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        const double* __restrict__ landing = X.peers.recv[X.me][X.parity];
        const unsigned long long* flags = X.peers.halo_flag[X.me];
        for (;;)
        {
            if (threadIdx.x == 0)
            {
                int done;
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(cnt + 2) : "memory");
                *slot = done >= X.n_recv ? X.n_recv : atomicAdd(cnt + 1, 1);
            }
            __syncthreads();
            const int n = *slot;
            __syncthreads();
            if (n >= X.n_recv) break;
            const halo_entry_dev_t e = X.recv[n];
            if (threadIdx.x == 0) bounded_wait_sys(flags + e.pad, X.counter, X.peers, X.me, e.pad);      // (gives up only when the run is called off)
            __syncthreads();
            const int cells = e.ni * e.nj;
            for (int k = threadIdx.x; k < 3 * cells; k += blockDim.x)
            {
                const int q = k / cells, c = k % cells;
                const int i = e.i0 + c / e.nj, j = e.j0 + c % e.nj;
                X.U[q * FS + (size_t(e.block) * N + i) * N + j] = __ldcg(landing + e.offset + k);     // written by a peer: not through L1
            }
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(cnt + 2, 1);
        }
        if (threadIdx.x == 0)
        {
            int done;
            unsigned n = 0;
            do
            {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(cnt + 2) : "memory");
                if ((++n & 1023u) == 0 && load_acquire_sys(X.peers.abort_word[X.me]) != 0) break;       // a CTA that holds strips gave up on its peer
            } while (done < X.n_recv);
            if (X.clock_words)
            {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                atomicAdd(X.clock_words + 2, t1 - t0);
                atomicAdd(X.clock_words + 3, 1ull);
            }
        }
        __syncthreads();
    }

    /** one input of the update phase, asked for where the call stands (volatile asm): NC = through ld.global.nc */
    template<int NC>
    __device__ __forceinline__ double ld_update(const double* p)
    {
        double v;
#ifdef M3B_NC_FREE
        if (NC) asm("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
#else
        if (NC) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
#endif
        else    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
        return v;
    }

    /** What the two low faces of a cell need from one of the three cells around them: primitives, the doubled PLM differences
     *  along the face normal, and the viscous combinations D1 = dx ux - dy uy, D2 = dx uy + dy ux (doubled, un-divided). */
    struct face_cell_t { double p[3], g[3], d1, d2; };

    /** cell (li, lj) of the tile as the x-faces see it (AXIS 0: differences along x) or as the y-faces see it (AXIS 1) */
    template<int AXIS, typename SMEM>
    __device__ __forceinline__ face_cell_t load_face_cell(const SMEM& T, const double (*P)[SX + 4][SY + 4], int li, int lj)
    {
        face_cell_t c;
        #pragma unroll
        for (int q = 0; q < 3; ++q) { c.p[q] = P[q][li + 2][lj + 2]; c.g[q] = T.G[3 * AXIS + q][li + 1][lj + 1]; }
        const double gx_vx = AXIS == 0 ? c.g[1] : T.G[1][li + 1][lj + 1], gx_vy = AXIS == 0 ? c.g[2] : T.G[2][li + 1][lj + 1];
        const double gy_vx = AXIS == 1 ? c.g[1] : T.G[4][li + 1][lj + 1], gy_vy = AXIS == 1 ? c.g[2] : T.G[5][li + 1][lj + 1];
        c.d1 = gx_vx - gy_vy;
        c.d2 = gx_vy + gy_vx;
        return c;
    }

    /** cs2, cs and the viscous coefficient at a face from the tabulated squared distances */
    template<bool FAST>
    __device__ __forceinline__ eos_face_t tma_eos(const model_t& model, const stage_t& S, const strip_consts_t& C, double cvis, double d1, double d2, double q2)
    {
        eos_face_t e;
        if (FAST) e = eos_face_fast(C, d1, d2, q2);
        else { const eos_t g = eos_from_distances<false>(model, S, d1, d2, q2); e.cs2 = g.cs2; e.cs = g.cs; e.mu_coef = cvis * g.nu; }
        return e;
    }

    /** face between the cells l (low side) and r along AXIS: intercell_flux_u (scheme.cpp:268-293) */
    template<int AXIS>
    __device__ __forceinline__ void tma_face(const eos_face_t& e, const face_cell_t& l, const face_cell_t& r, double F[3])
    {
        const prim_t L = {fma(l.g[0], 0.25, l.p[0]), fma(l.g[1], 0.25, l.p[1]), fma(l.g[2], 0.25, l.p[2])};
        const prim_t R = {fma(r.g[0], -0.25, r.p[0]), fma(r.g[1], -0.25, r.p[1]), fma(r.g[2], -0.25, r.p[2])};
        hlle_viscous_core<AXIS>(e.cs2, e.cs, e.mu_coef, L, R, l.d1 + r.d1, l.d2 + r.d2, F);
    }

    /** x-face between tile cells (li - 1, lj) and (li, lj), 0 <= li <= SX; yd[k] = T.y2c[k][lj] (tile-boundary faces: nothing to reuse) */
    template<bool FAST, typename SMEM>
    __device__ __forceinline__ void tma_x_face(const SMEM& T, const double (*P)[SX + 4][SY + 4], const model_t& model, const stage_t& S,
        const strip_consts_t& C, double cvis, const double yd[3], int li, int lj, double F[3])
    {
        const eos_face_t e = tma_eos<FAST>(model, S, C, cvis, T.x2v[0][li] + yd[0], T.x2v[1][li] + yd[1], T.x2v[2][li] + yd[2]);
        tma_face<0>(e, load_face_cell<0>(T, P, li - 1, lj), load_face_cell<0>(T, P, li, lj), F);
    }

    /** y-face between tile cells (li, lj - 1) and (li, lj), 0 <= lj <= SY; yd[k] = T.y2v[k][lj] */
    template<bool FAST, typename SMEM>
    __device__ __forceinline__ void tma_y_face(const SMEM& T, const double (*P)[SX + 4][SY + 4], const model_t& model, const stage_t& S,
        const strip_consts_t& C, double cvis, const double yd[3], int li, int lj, double F[3])
    {
        const eos_face_t e = tma_eos<FAST>(model, S, C, cvis, T.x2c[0][li] + yd[0], T.x2c[1][li] + yd[1], T.x2c[2][li] + yd[2]);
        tma_face<1>(e, load_face_cell<1>(T, P, li, lj - 1), load_face_cell<1>(T, P, li, lj), F);
    }

    /**
     * Both low faces of cell (li, lj): the cell itself is read once for the two of them, and the cell above comes from the
     * row before in registers (`up` in, this cell out) -- 17 shared-memory loads per cell instead of 32.
     */
    template<bool FAST, typename SMEM>
    __device__ __forceinline__ void tma_cell_faces(const SMEM& T, const double (*P)[SX + 4][SY + 4], const model_t& model, const stage_t& S,
        const strip_consts_t& C, double cvis, const double ydc[3], const double ydv[3], int li, int lj, face_cell_t& up, double Fx[3], double Fy[3])
    {
        const eos_face_t ex = tma_eos<FAST>(model, S, C, cvis, T.x2v[0][li] + ydc[0], T.x2v[1][li] + ydc[1], T.x2v[2][li] + ydc[2]);
        const eos_face_t ey = tma_eos<FAST>(model, S, C, cvis, T.x2c[0][li] + ydv[0], T.x2c[1][li] + ydv[1], T.x2c[2][li] + ydv[2]);
        // this cell for both faces: primitives once, all six differences once
        face_cell_t cx, cy;
        #pragma unroll
        for (int q = 0; q < 3; ++q) { cx.p[q] = cy.p[q] = P[q][li + 2][lj + 2]; cx.g[q] = T.G[q][li + 1][lj + 1]; cy.g[q] = T.G[3 + q][li + 1][lj + 1]; }
        cx.d1 = cy.d1 = cx.g[1] - cy.g[2];
        cx.d2 = cy.d2 = cx.g[2] + cy.g[1];
        tma_face<0>(ex, up, cx, Fx);
        tma_face<1>(ey, load_face_cell<1>(T, P, li, lj - 1), cy, Fy);
        up = cx;
    }

    /** what tma_rows needs to know about the tile */
    struct rows_args_t
    {
        double (*P)[SX + 4][SY + 4];
        double dt_over_h, cvis, yc, dy1, dy2;
        int b, i0, j0, N;
        size_t FS;
        const double* __restrict__ Uin; const double* __restrict__ Un; double* __restrict__ Uout;
        const double* __restrict__ U0; const double* __restrict__ BR;
        fail_dev_t* fail;
    };

    /**
     * Phases 2 + 3 of a tile for one warp: the faces of its strip of 4 rows and the update of its cells (block_update_u,
     * scheme.cpp:568-587), software-pipelined down the strip.  SINK: some cell of the tile lies within the sinks' reach
     * (warp-uniform, decided per tile); has_buffer: the buffer-zone rate is non-zero somewhere in the tile (else its inputs are not loaded).  Contains ONE __syncthreads.
     */
    template<bool FAST, int MODE, bool SINK, typename SMEM>
    __device__ __forceinline__ void tma_rows(SMEM& T, const model_t& model, const stage_t& S, const strip_consts_t& C, const rows_args_t& A,
        bool has_buffer, bool combine, bool compute_dt, strip_sums_t& sums, double& amax)
    {
        const int lane = threadIdx.x & 31, warp = logical_warp();
        const int li0 = STRIP * warp, lj = lane, N = A.N;
        const size_t FS = A.FS;
        const double yc = A.yc, dy1 = A.dy1, dy2 = A.dy2, cvis = A.cvis, dt_over_h = A.dt_over_h;
        const double ydc[3] = {T.y2c[0][lj], T.y2c[1][lj], T.y2c[2][lj]};      // x-faces and cell centres share the column's y-part
        const double ydv[3] = {T.y2v[0][lj], T.y2v[1][lj], T.y2v[2][lj]};
        const size_t c0 = (size_t(A.b) * N + (A.i0 + li0)) * N + (A.j0 + lj);
        unsigned negative = 0;          // bit r: the new density of strip row r is negative (its value waits in the cell's slot of P)

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
            // The cell's state reaches the source terms through a select on a flux of this row (true for every finite flux):
            // without it ptxas hoists the source terms to just behind the loads that deliver u, u0 and br, where the in-order
            // warp then sits on the long scoreboard with a whole row of independent face arithmetic queued behind it.
#ifdef M3B_NO_LATE
            const bool late = true;
#else
            const bool late = __double2hiint(FxHi[0]) != 0x7ff80001;
#endif
            const double us = late ? u[0] : 0.0, upx = late ? u[1] : 0.0, upy = late ? u[2] : 0.0;
            double acc[3], y1, y2;
            source_terms_strip<FAST>(model, C, x, yc, T.dxc[0][li], dy1, T.dxc[1][li], dy2, T.x2c[0][li] + ydc[0], T.x2c[1][li] + ydc[1],
                SINK, true, us, upx, upy, u0[0], u0[1], u0[2], br, acc, sums, y1, y2, T.sinks[warp]);

            double n0 = fma(-((FxHi[0] - FxLo[0]) + (hy[0] - FyLo[0])), dt_over_h, acc[0]);
            double n1 = fma(-((FxHi[1] - FxLo[1]) + (hy[1] - FyLo[1])), dt_over_h, acc[1]);
            double n2 = fma(-((FxHi[2] - FxLo[2]) + (hy[2] - FyLo[2])), dt_over_h, acc[2]);

#ifndef M3B_HOT_PATH_ONLY
            // validate_u (scheme.cpp:726-752) without a branch in the strip: remember the value, report after the last row
            const bool neg = __double2hiint(n0) < 0;
            if (neg) A.P[0][li + 2][lj + 2] = n0;           // (the cell's own sigma slot: nobody reads it any more)
            negative |= neg ? 1u << r : 0u;
#endif
            if (combine)
            {
                const double w = 1.0 - S.rk_b0;         // (the loaded un only meets the late value: nothing to hoist)
                n0 = fma(un[0], S.rk_b0, n0 * w);
                n1 = fma(un[1], S.rk_b0, n1 * w);
                n2 = fma(un[2], S.rk_b0, n2 * w);
            }
            A.Uout[c] = n0; A.Uout[FS + c] = n1; A.Uout[2 * FS + c] = n2;

            if (compute_dt)
                amax = dmax(amax, FAST ? max_wavespeed_fast(C, y1, y2, n0, n1, n2) : max_wavespeed<false>(model, S, x, yc, y1, y2, n0, n1, n2));
        };
        auto load_cell = [&] (int r, double* u, double* u0, double& br)
        {
            // volatile: issued here, a row of face arithmetic before their use
            const size_t c = c0 + size_t(r) * N;
            u[0] = ld_update<M3B_LD_SEL(0)>(A.Uin + c); u[1] = ld_update<M3B_LD_SEL(0)>(A.Uin + FS + c); u[2] = ld_update<M3B_LD_SEL(0)>(A.Uin + 2 * FS + c);
            br = 0.0; u0[0] = u0[1] = u0[2] = 0.0;
            if (has_buffer)
            {
                br = ld_update<M3B_LD_SEL(1)>(A.BR + c);
                u0[0] = ld_update<M3B_LD_SEL(1)>(A.U0 + c); u0[1] = ld_update<M3B_LD_SEL(1)>(A.U0 + FS + c); u0[2] = ld_update<M3B_LD_SEL(1)>(A.U0 + 2 * FS + c);
            }
        };
        // the step-start state for the RK combination only meets the finished update: one register set, asked for one row of
        // faces ahead of its use
        auto load_un = [&] (int r, double* un)
        {
            const size_t c = c0 + size_t(r) * N;
            un[0] = un[1] = un[2] = 0.0;
            if (combine) { un[0] = ld_update<M3B_LD_SEL(2)>(A.Un + c); un[1] = ld_update<M3B_LD_SEL(2)>(A.Un + FS + c); un[2] = ld_update<M3B_LD_SEL(2)>(A.Un + 2 * FS + c); }
        };

        // DEEP (3 CTAs per SM, 168 registers): the update inputs of row r are asked for BEFORE the faces of row r and used after
        // the faces of row r + 1 (two register sets, alternating): at least a row of face arithmetic lies between every load and
        // its use, also for the last row of the strip.  Otherwise (4 CTAs per SM, 128 registers): one set, asked for before the
        // faces of the NEXT row -- a row of faces between load and use -- and the last row's behind its own faces.
        constexpr bool DEEP = sizeof(T.P) > sizeof(T.P[0]);
        double u[2][3], u0[2][3], un[3], un_last[3], br[2];
        if (DEEP) load_cell(0, u[0], u0[0], br[0]);        // in flight during the prologue

        // prologue: tile-boundary faces (high-x row by warp 0, high-y column by half of warp 1), then the faces of strip
        // row 0, whose x-flux is also the high-x flux of the strip below
        if (warp == 0)
        {
            double F[3];
            tma_x_face<FAST>(T, A.P, model, S, C, cvis, ydc, SX, lj, F);
            T.XB[0][3][lj] = F[0]; T.XB[1][3][lj] = F[1]; T.XB[2][3][lj] = F[2];
        }
        else if (warp == 1 && lane < SX)
        {
            double F[3];
            const double ydhi[3] = {T.y2v[0][SY], T.y2v[1][SY], T.y2v[2][SY]};
            tma_y_face<FAST>(T, A.P, model, S, C, cvis, ydhi, lane, SY, F);
            T.YB[0][lane] = F[0]; T.YB[1][lane] = F[1]; T.YB[2][lane] = F[2];
        }
        // (the fluxes of consecutive rows alternate between two register sets: no copies at the end of a row)
        double Fx[2][3], Fy[2][3];
        face_cell_t up = load_face_cell<0>(T, A.P, li0 - 1, lj);        // the cell above the strip, then each row's cell for the row below
        tma_cell_faces<FAST>(T, A.P, model, S, C, cvis, ydc, ydv, li0, lj, up, Fx[0], Fy[0]);

        if (warp > 0) { T.XB[0][warp - 1][lj] = Fx[0][0]; T.XB[1][warp - 1][lj] = Fx[0][1]; T.XB[2][warp - 1][lj] = Fx[0][2]; }
        __syncthreads();

        #pragma unroll
        for (int r = 1; r < STRIP; ++r)
        {
            const int set = DEEP ? ((r - 1) & 1) : 0;
            if (DEEP) load_cell(r, u[r & 1], u0[r & 1], br[r & 1]); else load_cell(r - 1, u[0], u0[0], br[0]);
            load_un(r - 1, un);
            if (DEEP && r == STRIP - 1) load_un(r, un_last);     // (the last row has no faces of a next row to hide behind)
            if (! DEEP) up = load_face_cell<0>(T, A.P, li0 + r - 1, lj);      // (128 registers: re-read the cell above instead of carrying it)
            tma_cell_faces<FAST>(T, A.P, model, S, C, cvis, ydc, ydv, li0 + r, lj, up, Fx[r & 1], Fy[r & 1]);
            if (! DEEP && r == STRIP - 1) { load_cell(r, u[1], u0[1], br[1]); load_un(r, un_last); }       // behind them: the update of row r - 1
            update_cell(r - 1, u[set], u0[set], br[set], un, Fx[(r - 1) & 1], Fx[r & 1], Fy[(r - 1) & 1]);
        }
        {
            double FxHi[3];
            FxHi[0] = T.XB[0][warp][lj]; FxHi[1] = T.XB[1][warp][lj]; FxHi[2] = T.XB[2][warp][lj];
            update_cell(STRIP - 1, u[(STRIP - 1) & 1], u0[(STRIP - 1) & 1], br[(STRIP - 1) & 1], un_last, Fx[(STRIP - 1) & 1], FxHi, Fy[(STRIP - 1) & 1]);
        }
#ifndef M3B_HOT_PATH_ONLY
        if (negative)
        {
            for (int r = 0; r < STRIP; ++r)
                if (negative >> r & 1)
                {
                    const double v = A.P[0][li0 + r + 2][lj + 2];
                    if (v < 0.0) report_negative(A.fail, A.b, (A.i0 + li0 + r) * N + A.j0 + lj, v);      // (not -0.0)
                }
        }
#endif
    }

    template<int MIN_CTAS, int NB, bool FAST, int MODE, int NBUF = (MIN_CTAS >= 4 ? 1 : 2)>
    __global__ void __launch_bounds__(STRIP_THREADS, MIN_CTAS) stage_tma(
        mesh_dev_t mesh, model_t model, const stage_t* __restrict__ stage_ptr, const tile_info_t* __restrict__ tile_info, int num_tiles,
        const double* __restrict__ Uin, const double* __restrict__ Un, double* __restrict__ Uout, double* partials, double* jump_partials,
        fail_dev_t* fail, fused_exchange_t X)
    {
        extern __shared__ __align__(128) unsigned char smem_raw[];
        using SMEM = tma_smem_t<NBUF>;
        SMEM& T = *reinterpret_cast<SMEM*>(smem_raw);

        const stage_t S = *stage_ptr;       // written by the host or by prepare_next of the step before
        // MODE 1 / 2: first / last stage of an RK2 step with adaptive dt, flags known at compile time; 0: read them from S
        const bool combine = MODE == 0 ? S.combine != 0 : MODE == 2, compute_dt = MODE == 0 ? S.compute_dt != 0 : MODE == 2;
        const int N = NB ? NB : mesh.N;
        const int tiles_y = N / SY;
        const size_t FS = mesh.FS;
        const int lane = threadIdx.x & 31, warp = logical_warp();
        const strip_consts_t C = {S.m1 * model.inv_mach2, S.m2 * model.inv_mach2, -S.m1, -S.m2, 2.0 * S.theta, S.dt};
        const double* __restrict__ U0 = mesh.U0;
        const double* __restrict__ BR = mesh.br;

        // ordinal k of this CTA's tiles <-> tile blockIdx.x + k gridDim.x; record ring slot k % 3, P half k & 1
        int tile = blockIdx.x;
        if (X.enabled) exchange_push(X, FS, N, &T.unpack_slot);       // (every CTA of the grid takes part, also one without tiles)
        if (tile >= num_tiles) return;
        bool ghosts_ready = ! X.enabled;
        if (threadIdx.x < 6)
        {
            const int which = threadIdx.x / 3, part = threadIdx.x % 3;
            const int tl = tile + which * int(gridDim.x);
            if (tl < num_tiles) reinterpret_cast<int4*>(&T.info[which])[part] = __ldg(reinterpret_cast<const int4*>(tile_info + tl) + part);
        }
        __syncthreads();
        if (! ghosts_ready && tile >= mesh.first_wait_cta) { exchange_unpack(X, FS, N, &T.unpack_slot); ghosts_ready = true; }
        stage_tile_async(T, T.info[0], Uin, FS, T.info[0].flags >> TILE_POS_SHIFT, 0, N, tiles_y);
        if (warp == 0)
        {
            const int b0 = T.info[0].b, t0 = T.info[0].flags >> TILE_POS_SHIFT;
            const double* xvg = mesh.xv + size_t(b0) * (N + 1) + (t0 / tiles_y) * SX;
            const double* yvg = mesh.yv + size_t(b0) * (N + 1) + (t0 % tiles_y) * SY;
            if (lane <= SX) T.cx[0][lane] = __ldg(xvg + lane);
            T.cy[0][lane] = __ldg(yvg + lane);
            if (lane == 0) T.cy[0][SY] = __ldg(yvg + SY);
            if (lane == 1) { T.hh[0][0] = __ldg(mesh.spacing + b0); T.hh[0][1] = __ldg(mesh.inv_spacing + b0); }
        }
        __syncthreads();

        for (int k = 0; tile < num_tiles; ++k, tile += gridDim.x)
        {
            const int buf = NBUF == 2 ? (k & 1) : 0, nbuf = NBUF == 2 ? (buf ^ 1) : 0;       // P half of this tile / of the next; cx, cy, hh alternate either way
            const int cb = k & 1;
            const int next = tile + int(gridDim.x);
            const bool has_next = next < num_tiles;
            const tile_info_t& tin = T.info[(k + 1) % 3];
            double nx = 0.0, ny = 0.0, ny32 = 0.0;      // warp 0: vertex coordinates of the next tile, stored when they have arrived
            int4 rec = make_int4(0, 0, 0, 0);           // threads 32-34: the record of the tile after the next
            // the other half of P was released by the barrier that ended tile k - 1: refill it while this tile is computed
            if (has_next)
            {
                if (! ghosts_ready && next >= mesh.first_wait_cta) { exchange_unpack(X, FS, N, &T.unpack_slot); ghosts_ready = true; }
                const int tn = tin.flags >> TILE_POS_SHIFT;
                if (NBUF == 2) stage_tile_async(T, tin, Uin, FS, tn, nbuf, N, tiles_y);
                else stage_tile_async<true>(T, tin, Uin, FS, tn, 0, N, tiles_y);
                if (warp == 0)
                {
                    const double* xvg = mesh.xv + size_t(tin.b) * (N + 1) + (tn / tiles_y) * SX;
                    const double* yvg = mesh.yv + size_t(tin.b) * (N + 1) + (tn % tiles_y) * SY;
                    nx = lane <= SX ? __ldg(xvg + lane) : 0.0;
                    ny = __ldg(yvg + lane);
                    ny32 = lane == 0 ? __ldg(yvg + SY) : (lane == 1 ? __ldg(mesh.spacing + tin.b) : (lane == 2 ? __ldg(mesh.inv_spacing + tin.b) : 0.0));
                }
                else if (warp == 1 && lane < 3 && next + int(gridDim.x) < num_tiles)
                    rec = __ldg(reinterpret_cast<const int4*>(tile_info + next + gridDim.x) + lane);
            }
            else if (NBUF == 2) cp_async_commit();     // (an empty group keeps the wait below uniform)

            const int b = T.info[k % 3].b, flags = T.info[k % 3].flags;
            const int t = flags >> TILE_POS_SHIFT;
            const int i0 = (t / tiles_y) * SX, j0 = (t % tiles_y) * SY;
            const bool has_buffer = flags & 1;
            const double h = T.hh[cb][0], inv_h = T.hh[cb][1];
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
                        const double xv = T.cx[cb][lane];
                        #pragma unroll
                        for (int q = 0; q < 3; ++q) { const double d = (xv - bx[q]) * (q == 2 ? cf : 1.0); T.x2v[q][lane] = d * d; }
                    }
                    if (lane < SX)
                    {
                        const double xc = 0.5 * (T.cx[cb][lane] + T.cx[cb][lane + 1]);
                        T.xc[lane] = xc;
                        T.dxc[0][lane] = xc - bx[0];
                        T.dxc[1][lane] = xc - bx[1];
                        #pragma unroll
                        for (int q = 0; q < 3; ++q) { const double d = (xc - bx[q]) * (q == 2 ? cf : 1.0); T.x2c[q][lane] = d * d; }
                    }
                }
                else if (warp == 2)
                {
                    const double yv = T.cy[cb][lane], yc = 0.5 * (yv + T.cy[cb][lane + 1]);
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
                    const double yv = T.cy[cb][SY];
                    #pragma unroll
                    for (int q = 0; q < 3; ++q) { const double dv = (yv - by[q]) * (q == 2 ? cf : 1.0); T.y2v[q][SY] = fma(dv, dv, soft[q]); }
                    // does any cell of the tile lie within the sinks' reach (a2 = dr^2 / (2 s^2) < SINK_REACH_A2)?  distance of each body to the tile's rectangle
                    const double xlo = T.cx[cb][0], xhi = T.cx[cb][SX], ylo = T.cy[cb][0], yhi = yv;
                    bool near = false;
                    #pragma unroll
                    for (int q = 0; q < 2; ++q)
                    {
                        const double ddx = dmax(dmax(xlo - bx[q], bx[q] - xhi), 0.0), ddy = dmax(dmax(ylo - by[q], by[q] - yhi), 0.0);
                        near = near || (ddx * ddx + ddy * ddy) * model.sink_inv_2s2 < SINK_REACH_A2;
                    }
                    T.near_sink = near;
                }
            }

            // ------------------------------------------------------------------ phase 0: this thread's chunks of the tile, primitives in place
            // (all but the newest group: the chunks asked for during tile k - 1)
            if (NBUF == 2) asm volatile("cp.async.wait_group 1;" ::: "memory"); else cp_async_wait_all();
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

                #pragma unroll (PLM_UNROLL)
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
                    if (lane <= SX) T.cx[cb ^ 1][lane] = nx;
                    T.cy[cb ^ 1][lane] = ny;
                    if (lane == 0) T.cy[cb ^ 1][SY] = ny32;
                    if (lane == 1 || lane == 2) T.hh[cb ^ 1][lane - 1] = ny32;
                }
                else if (warp == 1 && lane < 3 && has_next && next + int(gridDim.x) < num_tiles)
                    reinterpret_cast<int4*>(&T.info[(k + 2) % 3])[lane] = rec;
            }
            __syncthreads();

            // ------------------------------------------------------------------ phases 2 + 3: faces, update
            strip_sums_t sums = {{0.0, 0.0}, {0.0, 0.0}, 0.0, 0.0, 0.0};
            double amax = 0.0;                              // largest signal speed of the updated cells
            const double yc = 0.5 * (T.cy[cb][lane] + T.cy[cb][lane + 1]);
            const double dy1 = yc - S.y1, dy2 = yc - S.y2;
            {
                const rows_args_t A = {P, S.dt * inv_h, cvis, yc, dy1, dy2, b, i0, j0, N, FS, Uin, Un, Uout, U0, BR, fail};
                // one straight-line variant per (sink within reach, buffer zone) so that the common path has no branch: ptxas
                // schedules inside basic blocks, and only an unbroken strip lets it overlap a row's update with the next row's faces
#ifdef M3B_HOT_PATH_ONLY
                tma_rows<FAST, MODE, false>(T, model, S, C, A, has_buffer, combine, compute_dt, sums, amax);
#else
                // (the buffer terms are not a variant: without a rate they add exact zeros, and a second copy of the strip would
                // have to share the instruction cache with the first)
                if (T.near_sink != 0) tma_rows<FAST, MODE, true>(T, model, S, C, A, has_buffer, combine, compute_dt, sums, amax);
                else                  tma_rows<FAST, MODE, false>(T, model, S, C, A, has_buffer, combine, compute_dt, sums, amax);
#endif
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
                const double v1 = warp_sum8(v8, lane);
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
                double* row = ((flags & TILE_JUMP_ROWS) ? jump_partials : partials) + size_t(T.info[k % 3].row) * ROW;
                const double a = T.red[0][q], bq = T.red[1][q], cq = T.red[2][q], d = T.red[3][q];
                // min over cells of h / wavespeed = h / max wavespeed: one division per tile
                row[q] = q == NUM_SUMS ? (compute_dt ? h / dmax(dmax(a, bq), dmax(cq, d)) : 1e300) : ((a + bq) + (cq + d)) * (h * h);
            }
            // one buffer: only now is P free for the next tile (its lines were pulled into L2 a tile ago)
            if (NBUF == 1 && has_next) stage_tile_async(T, tin, Uin, FS, tin.flags >> TILE_POS_SHIFT, 0, N, tiles_y);
        }
    }
}}}
