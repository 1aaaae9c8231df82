/**
 * finish_kernels.cuh -- what runs around the stage kernels: maximum_timestep without a fused estimate, the body positions of
 * the coming stages (prepare_positions), the fixed-order fold of the stage kernels' partial rows with the block-wise work
 * integrals, the cross-rank exchange of the stage results and the next step's time and dt (finish_stage, finish_stage_cluster,
 * peer_prepare, prepare_next), and the small kernels of the products and of the C ABI's state layout (disk totals, diagnostic
 * fields, permute, axpby).  Included by kernels.cu only (one translation unit with the launch logic).
 */
#pragma once
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

}
