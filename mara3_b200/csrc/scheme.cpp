#include <cstdlib>
#include "scheme.hpp"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <numeric>
#include "iso2d_sums.hpp"

using namespace m3b;

void solution_t::get_scalars(double o[num_scalars]) const
{
    o[0] = time; o[1] = iteration_num; o[2] = iteration_den;
    for (int k = 0; k < 2; ++k)
    {
        o[3 + k] = mass_accreted_on[k];
        o[5 + k] = angular_momentum_accreted_on[k];
        o[7 + k] = integrated_torque_on[k];
        o[9 + k] = work_done_on[k];
    }
    o[11] = mass_ejected;
    o[12] = angular_momentum_ejected;
    auto put = [o] (int at, const elements_t& e)
    {
        const double v[10] = {e.pomega, e.tau, e.cm_position_x, e.cm_position_y, e.cm_velocity_x, e.cm_velocity_y,
                              e.separation, e.total_mass, e.mass_ratio, e.eccentricity};
        for (int k = 0; k < 10; ++k) o[at + k] = v[k];
    };
    put(13, orbital_elements_acc);
    put(23, orbital_elements_grav);
    put(33, orbital_elements);
}

void solution_t::set_scalars(const double o[num_scalars])
{
    time = o[0]; iteration_num = int(o[1]); iteration_den = int(o[2]);
    for (int k = 0; k < 2; ++k)
    {
        mass_accreted_on[k] = o[3 + k];
        angular_momentum_accreted_on[k] = o[5 + k];
        integrated_torque_on[k] = o[7 + k];
        work_done_on[k] = o[9 + k];
    }
    mass_ejected = o[11];
    angular_momentum_ejected = o[12];
    auto get = [o] (int at)
    {
        return elements_t{o[at], o[at + 1], o[at + 2], o[at + 3], o[at + 4], o[at + 5], o[at + 6], o[at + 7], o[at + 8], o[at + 9]};
    };
    orbital_elements_acc = get(13);
    orbital_elements_grav = get(23);
    orbital_elements = get(33);
}

binary_solver_t::binary_solver_t(const config_t& run_config, int device, bool general_only, bool tiled_kernel,
                                 int rank, int nranks, const unsigned char* nccl_unique_id)
: config(run_config)
, data(create_solver_data(run_config, rank, nranks))
{
    // set_scheme_globals (scheme.cpp:42-49) rejects threaded <= 0; the key is otherwise unused here
    if (config.get_int("threaded") <= 0)
        throw std::invalid_argument("runtime option 'threaded' for number of threads must be > 0");
    if (data.rk_order != 1 && data.rk_order != 2)
        throw std::invalid_argument("binary::next_solution");
    if (device >= 0)
    {
        // conserve_linear_p = 0 (advance_q, scheme.cpp:906-1020): the state is conserved_q and every block takes the any-tree kernels
        // (regular blocks whose size is a multiple of 32 take the strip kernel's QMODE variant; M3B_Q_STRIP=0: any-tree kernels only)
        const char* qs = std::getenv("M3B_Q_STRIP");
        const bool q_strip = data.block_size % 32 == 0 && ! tiled_kernel && ! (qs && std::atoi(qs) == 0);
        gpu = std::make_unique<device_solver_t>(data, device, general_only || (! data.conserve_linear_p && ! q_strip), tiled_kernel);
        if (nranks > 1)
        {
            if (! nccl_unique_id) throw std::invalid_argument("a multi-rank solver needs the NCCL unique id of the job");
            comm = std::make_unique<communicator_t>(rank, nranks, nccl_unique_id);      // on the device the solver selected
            gpu->set_communicator(comm.get());
            auto maxima = gpu->all_gather_scalar(data.max_velocity_local);
            data.set_max_velocity(*std::max_element(maxima.begin(), maxima.end()));
        }
        scratch1 = new_field();
        scratch2 = new_field();
    }
}

device_solver_t& binary_solver_t::device()
{
    // host-only solvers (mesh queries, CPU tests) have no device context; there is no CPU compute path
    if (! gpu) throw std::runtime_error("mara3_b200: this solver was created without a CUDA device; compute calls need one");
    return *gpu;
}

std::shared_ptr<device_field_t> binary_solver_t::new_field()
{
    // fields released by solutions go back to a small pool instead of cudaFree: value-semantics callers
    // (m3b_advance_host, clone) would otherwise pay a cudaMalloc + cudaFree per call
    device_field_t* raw = nullptr;
    if (! pool->free.empty())
    {
        raw = pool->free.back().release();
        pool->free.pop_back();
    }
    else
    {
        raw = new device_field_t(device().state_doubles(), device().device());
    }
    auto keep = pool;
    return std::shared_ptr<device_field_t>(raw, [keep] (device_field_t* f)
    {
        if (keep->free.size() < 8) keep->free.emplace_back(f); else delete f;
    });
}

solution_t binary_solver_t::create_solution()
{
    auto s = solution_t();
    s.conserved_u = new_field();
    device().load_initial(*s.conserved_u);
    s.orbital_elements = data.initial_elements;
    return s;
}

solution_t binary_solver_t::clone(const solution_t& s)
{
    auto r = s;
    r.conserved_u = new_field();
    device().copy(*s.conserved_u, *r.conserved_u);
    return r;
}

double binary_solver_t::maximum_timestep(const solution_t& s)
{
    auto bodies = two_body_state(s.orbital_elements, s.time);
    device().launch_max_timestep(*s.conserved_u, s.time, bodies, 7);
    device().gather_results();
    device().sync();
    return device().stage_result(7).dt_min;
}

stage_inputs_t binary_solver_t::stage_inputs(const solution_t& in, double dt, bool safe_mode) const
{
    auto inputs = stage_inputs_t();
    inputs.time = in.time;
    inputs.dt = dt;
    inputs.theta = safe_mode ? 0.0 : data.plm_theta;                    // scheme.cpp:792
    inputs.bodies = two_body_state(in.orbital_elements, in.time);       // scheme.cpp:814
    return inputs;
}

/**
 * Host part of advance_u after the block updates: totals (scheme.cpp:390-408, 829-830),
 * accretion / gravity kicks and orbital elements (:832-885), the new scalars (:889-903).
 */
status_t binary_solver_t::bookkeeping(const solution_t& in, const stage_result_t& r, const two_body_t& bodies, double dt, solution_t& out)
{
    using namespace sums;
    const point_mass_t* body[2] = {&bodies.body1, &bodies.body2};
    double mass_acc[2], lz_acc[2], px_acc[2], py_acc[2], torque[2], fx[2], fy[2], work[2];

    for (int k = 0; k < 2; ++k)
    {
        mass_acc[k] = +r.sums[ACC_MASS + k] * dt;
        px_acc[k]   = +r.sums[ACC_PX + k] * dt;
        py_acc[k]   = +r.sums[ACC_PY + k] * dt;
        lz_acc[k]   = +r.sums[ACC_LZ + k] * dt;
        fx[k]       = -r.sums[GRV_FX + k] * dt;
        fy[k]       = -r.sums[GRV_FY + k] * dt;
        torque[k]   = -r.sums[GRV_TQ + k] * dt;

        // evaluated per block on the device, then summed (scheme.cpp:407-408); source_terms_q does not set it (:446-463)
        work[k]     = data.conserve_linear_p ? r.work[k] : 0.0;
    }
    double mass_ejected = -r.sums[BUF_M] * dt;
    double lz_ejected   = -r.sums[BUF_L] * dt;

    two_body_t acc = bodies, grv = bodies;
    point_mass_t* acc_body[2] = {&acc.body1, &acc.body2};
    point_mass_t* grv_body[2] = {&grv.body1, &grv.body2};

    for (int k = 0; k < 2; ++k)
    {
        double M = body[k]->mass;
        acc_body[k]->mass = M + mass_acc[k];
        if (! data.no_accretion_force)
        {
            acc_body[k]->vx = (M * body[k]->vx + px_acc[k]) / (M + mass_acc[k]);
            acc_body[k]->vy = (M * body[k]->vy + py_acc[k]) / (M + mass_acc[k]);
        }
        grv_body[k]->vx = body[k]->vx + fx[k] / M;
        grv_body[k]->vy = body[k]->vy + fy[k] / M;
    }
    elements_t E0 = in.orbital_elements, E_acc, E_grv;

    if (! orbital_elements(acc, in.time, E_acc) || ! orbital_elements(grv, in.time, E_grv))
    {
        error = "mara::compute_orbital_elements (two_body_state does not correspond to a bound orbit)";
        return status_unbound_orbit;
    }
    double live = in.time > data.begin_live_binary ? 1.0 : 0.0;      // compared in code units (scheme.cpp:882)
    auto d_acc = elements_diff(E0, E_acc);
    auto d_grv = elements_diff(E0, E_grv);
    auto conserved = out.conserved_u;

    out = in;
    out.conserved_u = conserved;
    out.time = in.time + dt;
    out.iteration_num = in.iteration_num + in.iteration_den;
    for (int k = 0; k < 2; ++k)
    {
        out.mass_accreted_on[k]             = in.mass_accreted_on[k] + mass_acc[k];
        out.angular_momentum_accreted_on[k] = in.angular_momentum_accreted_on[k] + lz_acc[k];
        out.integrated_torque_on[k]         = in.integrated_torque_on[k] + torque[k];
        out.work_done_on[k]                 = in.work_done_on[k] + work[k];
    }
    out.mass_ejected             = in.mass_ejected + mass_ejected;
    out.angular_momentum_ejected = in.angular_momentum_ejected + lz_ejected;
    out.orbital_elements_acc  = in.orbital_elements_acc + d_acc;
    out.orbital_elements_grav = in.orbital_elements_grav + d_grv;
    out.orbital_elements      = in.orbital_elements + (d_acc + d_grv + elements_cm_drift(E0, dt)) * live;
    return status_ok;
}

void binary_solver_t::record_offenders(int slot)
{
    messages.clear();
    const int N = data.block_size;
    auto n = device().local_num_negative(slot);

    for (const auto& o : device().offenders(slot))
    {
        int i = o.cell / N, j = o.cell % N;
        double x = (data.xv[std::size_t(o.block) * (N + 1) + i] + data.xv[std::size_t(o.block) * (N + 1) + i + 1]) * 0.5;
        double y = (data.yv[std::size_t(o.block) * (N + 1) + j] + data.yv[std::size_t(o.block) * (N + 1) + j + 1]) * 0.5;
        char line[160];
        std::snprintf(line, sizeof(line), "negative density %3.2e (at position [%+3.2lf %+3.2lf])", o.sigma, x, y);   // scheme.cpp:741
        messages.push_back(line);
    }
    if (n > unsigned(device_solver_t::max_offenders))
        messages.push_back("... (" + std::to_string(n - device_solver_t::max_offenders) + " more cells)");
    if (! quiet) for (const auto& m : messages) std::printf("%s\n", m.c_str());
}

status_t binary_solver_t::advance(const solution_t& in, double dt, bool safe_mode, solution_t& out)
{
    if (! out.conserved_u || out.conserved_u == in.conserved_u) out.conserved_u = new_field();

    auto inputs = stage_inputs(in, dt, safe_mode);
    device().launch_stage(*in.conserved_u, nullptr, *out.conserved_u, inputs, 0);
    device().gather_results();
    device().sync();
    const auto r = device().stage_result(0);
    auto st = bookkeeping(in, r, inputs.bodies, dt, out);
    if (st != status_ok) return st;

    if (r.num_negative)
    {
        record_offenders(0);
        error = "negative density in updated state";
        return status_negative_density;
    }
    return status_ok;
}

/** Scalar part of s0 * b0 + s2 * (1 - b0) (scheme.cpp:1033-1069); b0 = p / 2 in every use. */
static solution_t combine_scalars(const solution_t& s0, const solution_t& s2, double b0)
{
    double b1 = 1.0 - b0;
    auto r = solution_t();
    r.time = s0.time * b0 + s2.time * b1;

    // rational arithmetic of the iteration counter (core_rational.hpp:57-66, 96-104, 184-196)
    long q = 2, p0 = std::lround(b0 * q), p1 = q - p0;
    long num = long(s0.iteration_num) * p0 * s2.iteration_den + long(s2.iteration_num) * p1 * s0.iteration_den;
    long den = long(s0.iteration_den) * s2.iteration_den * q;
    long g = std::gcd(std::labs(num), std::labs(den));
    if (g == 0) g = 1;
    r.iteration_num = int(num / g);
    r.iteration_den = int(den / g);

    for (int k = 0; k < 2; ++k)
    {
        r.mass_accreted_on[k]             = s0.mass_accreted_on[k] * b0 + s2.mass_accreted_on[k] * b1;
        r.angular_momentum_accreted_on[k] = s0.angular_momentum_accreted_on[k] * b0 + s2.angular_momentum_accreted_on[k] * b1;
        r.integrated_torque_on[k]         = s0.integrated_torque_on[k] * b0 + s2.integrated_torque_on[k] * b1;
        r.work_done_on[k]                 = s0.work_done_on[k] * b0 + s2.work_done_on[k] * b1;
    }
    r.mass_ejected             = s0.mass_ejected * b0 + s2.mass_ejected * b1;
    r.angular_momentum_ejected = s0.angular_momentum_ejected * b0 + s2.angular_momentum_ejected * b1;
    r.orbital_elements_acc  = s0.orbital_elements_acc * b0 + s2.orbital_elements_acc * b1;
    r.orbital_elements_grav = s0.orbital_elements_grav * b0 + s2.orbital_elements_grav * b1;
    r.orbital_elements      = s0.orbital_elements * b0 + s2.orbital_elements * b1;
    return r;
}

void binary_solver_t::combine(const solution_t& s0, const solution_t& s2, double b0, solution_t& out)
{
    auto field = out.conserved_u && out.conserved_u != s0.conserved_u && out.conserved_u != s2.conserved_u ? out.conserved_u : new_field();
    device().combine(*s0.conserved_u, b0, *s2.conserved_u, 1.0 - b0, *field);
    out = combine_scalars(s0, s2, b0);
    out.conserved_u = field;
}

/**
 * One full step with the RK combination fused into the last stage.  On success `s`
 * holds the new solution (its field is swapped with internal scratch); on failure `s`
 * is untouched, as the reference's value semantics guarantee.
 */
status_t binary_solver_t::try_step(solution_t& s, double dt, bool safe_mode)
{
    auto& A = s.conserved_u;
    auto inputs1 = stage_inputs(s, dt, safe_mode);
    auto s1 = solution_t();
    auto s2 = solution_t();
    bool live_possible = s.time + 2 * dt > data.begin_live_binary;     // elements may change between the stages

    if (data.rk_order == 1)
    {
        inputs1.compute_dt = ! data.fixed_dt && ! live_possible;
        device().launch_stage(*A, nullptr, *scratch1, inputs1, 0);
        device().gather_results();
        device().sync();
        s1.conserved_u = scratch1;
        auto st = bookkeeping(s, device().stage_result(0), inputs1.bodies, dt, s1);
        if (st != status_ok) return st;
        if (device().stage_result(0).num_negative) { record_offenders(0); return status_negative_density; }
        std::swap(scratch1, A);
        auto field = A;
        s = s1;
        s.conserved_u = field;
        if (inputs1.compute_dt) { dt_cache_field = A.get(); dt_cache_time = s.time; dt_cache_value = device().stage_result(0).dt_min; }
        return status_ok;
    }

    // stage 1: A -> scratch1
    device().launch_stage(*A, nullptr, *scratch1, inputs1, 0);
    s1.conserved_u = scratch1;

    if (live_possible)
    {
        device().gather_results();
        device().sync();
        auto st = bookkeeping(s, device().stage_result(0), inputs1.bodies, dt, s1);
        if (st != status_ok) return st;
        if (device().stage_result(0).num_negative) { record_offenders(0); return status_negative_density; }
    }
    else
    {
        // orbital elements cannot change: the stage-2 inputs are known without waiting for stage 1
        s1.time = s.time + dt;
        s1.orbital_elements = s.orbital_elements;
    }
    // stage 2: scratch1 -> scratch2, fused with  s0 * 1/2 + s2 * 1/2  and the next CFL estimate
    auto inputs2 = stage_inputs(s1, dt, safe_mode);
    inputs2.combine = true;
    inputs2.rk_b0 = 0.5;
    inputs2.compute_dt = ! data.fixed_dt && ! live_possible;
    device().launch_stage(*scratch1, A.get(), *scratch2, inputs2, 1);
    device().gather_results();
    device().sync();

    if (! live_possible)
    {
        auto st = bookkeeping(s, device().stage_result(0), inputs1.bodies, dt, s1);
        if (st != status_ok) return st;
        if (device().stage_result(0).num_negative) { record_offenders(0); return status_negative_density; }
    }
    s2.conserved_u = scratch2;
    auto st = bookkeeping(s1, device().stage_result(1), inputs2.bodies, dt, s2);
    if (st != status_ok) return st;
    if (device().stage_result(1).num_negative) { record_offenders(1); return status_negative_density; }

    // scalars: s0 * b0 + s2 * (1 - b0); the field was already combined on the device
    auto result = combine_scalars(s, s2, 0.5);
    std::swap(scratch2, A);         // A now holds U^{n+1}; the old U^n becomes scratch
    auto field = A;
    s = result;
    s.conserved_u = field;

    if (inputs2.compute_dt)
    {
        dt_cache_field = A.get();
        dt_cache_time = s.time;
        dt_cache_value = device().stage_result(1).dt_min;
    }
    return status_ok;
}

bool binary_solver_t::can_pipeline(const solution_t& s, double dt) const
{
    // the device computes the next step's inputs from constant orbital elements: stay well clear of the
    // time at which the binary goes live (scheme.cpp:882)
    return pipelining && data.rk_order == 2 && s.time + 1000.0 * dt < data.begin_live_binary;
}

void binary_solver_t::drop_speculation()
{
    if (speculation.valid)
    {
        device().sync();            // its kernels may still be running on the buffers we are about to release
        speculation = speculation_t();
    }
}

static bool same_elements(const elements_t& a, const elements_t& b)
{
    return a.pomega == b.pomega && a.tau == b.tau && a.cm_position_x == b.cm_position_x && a.cm_position_y == b.cm_position_y
        && a.cm_velocity_x == b.cm_velocity_x && a.cm_velocity_y == b.cm_velocity_y && a.separation == b.separation
        && a.total_mass == b.total_mass && a.mass_ratio == b.mass_ratio && a.eccentricity == b.eccentricity;
}

void binary_solver_t::launch_pipelined(const std::shared_ptr<device_field_t>& in, const std::shared_ptr<device_field_t>& out, int parity, const elements_t& elements)
{
    device().launch_step_async(*in, *scratch1, *out, parity, elements, data.cfl_number, data.recommended_time_step,
                               data.plm_theta, data.fixed_dt);
}

/** Wait for a queued step, do the host bookkeeping of its two stages, and commit it to `s`. */
status_t binary_solver_t::finish_pipelined(solution_t& s, const speculation_t& step)
{
    device().wait_step(step.parity);
    const auto r1 = device().async_result(step.parity, 0);
    const auto r2 = device().async_result(step.parity, 1);
    const double dt = step.dt;

    auto inputs1 = stage_inputs(s, dt, false);
    auto s1 = solution_t(), s2 = solution_t();
    auto st = bookkeeping(s, r1, inputs1.bodies, dt, s1);
    if (st != status_ok) return st;
    if (r1.num_negative) { record_offenders(device().async_slot(step.parity, 0)); return status_negative_density; }
    auto inputs2 = stage_inputs(s1, dt, false);
    st = bookkeeping(s1, r2, inputs2.bodies, dt, s2);
    if (st != status_ok) return st;
    if (r2.num_negative) { record_offenders(device().async_slot(step.parity, 1)); return status_negative_density; }

    auto result = combine_scalars(s, s2, 0.5);
    result.conserved_u = step.out;
    s = result;
    dt_cache_field = s.conserved_u.get();
    dt_cache_time = s.time;
    dt_cache_value = r2.dt_min;
    return status_ok;
}

status_t binary_solver_t::next_solution(solution_t& s, double* dt_used, bool* fell_back, bool speculate)
{
    if (fell_back) *fell_back = false;

    // ---- a step queued earlier for exactly this state?
    bool queued = speculation.valid && speculation.in == s.conserved_u && speculation.time == s.time
        && same_elements(speculation.elements, s.orbital_elements);
    if (speculation.valid && ! queued) drop_speculation();

    double dt;
    if (queued)
    {
        dt = speculation.dt;
    }
    else if (data.fixed_dt)
    {
        dt = data.recommended_time_step;
    }
    else if (dt_cache_field == s.conserved_u.get() && dt_cache_time == s.time)
    {
        dt = data.cfl_number * dt_cache_value;      // produced by the previous step's last stage
    }
    else
    {
        dt = data.cfl_number * maximum_timestep(s);
    }
    dt_cache_field = nullptr;
    auto st = status_ok;

    if (queued || can_pipeline(s, dt))
    {
        auto step = speculation;
        speculation = speculation_t();

        if (! queued)
        {
            // start of a pipeline: the host provides the inputs of this step
            step.valid = true;
            step.in = s.conserved_u;
            step.out = new_field();
            step.time = s.time;
            step.dt = dt;
            step.parity = 0;
            auto inputs1 = stage_inputs(s, dt, false);
            auto after1 = s;
            after1.time = s.time + dt;
            auto inputs2 = stage_inputs(after1, dt, false);
            inputs2.combine = true;
            inputs2.rk_b0 = 0.5;
            inputs2.compute_dt = ! data.fixed_dt;
            device().upload_step_inputs(step.parity, inputs1, inputs2);
            launch_pipelined(step.in, step.out, step.parity, s.orbital_elements);
        }
        // queue the step after this one before waiting: its dt and body positions are already on the device
        if (speculate && can_pipeline(s, dt))
        {
            speculation.valid = true;
            speculation.in = step.out;
            speculation.out = new_field();
            speculation.parity = 1 - step.parity;
            speculation.elements = s.orbital_elements;      // constant while the binary is not live
            launch_pipelined(speculation.in, speculation.out, speculation.parity, s.orbital_elements);
        }
        st = finish_pipelined(s, step);

        if (st == status_ok && speculation.valid)
        {
            speculation.time = s.time;
            speculation.dt = data.fixed_dt ? data.recommended_time_step : data.cfl_number * dt_cache_value;
        }
        if (st != status_ok) drop_speculation();        // it started from a state that is being thrown away
    }
    else
    {
        st = try_step(s, dt, false);
    }

    if (st == status_negative_density || st == status_unbound_orbit)
    {
        // the reference catches any std::exception, prints it, and redoes the step in safe mode
        // (subprog_binary.cpp:285-292)
        if (st == status_negative_density) error = "negative density in updated state";
        if (! quiet) std::cout << error << std::endl;
        dt *= 0.1;
        if (fell_back) *fell_back = true;
        dt_cache_field = nullptr;
        st = try_step(s, dt, true);
    }
    if (dt_used) *dt_used = dt;
    return st;
}
