/**
 * two_body.hpp -- Kepler two-body model used by the `binary` hot path.
 *
 * Same quantities and formulas as the reference's model_two_body.hpp
 * (compute_two_body_state :168-281, compute_orbital_elements :295-402, diff /
 * diff_cm :492-529), written as plain host+device inline functions so the
 * per-step bookkeeping can run on the host or inside a device-side step
 * epilogue.  O(1) work per RK stage.
 */
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define M3B_HD __host__ __device__ inline
#else
#define M3B_HD inline
#endif

namespace m3b
{
    /** full_orbital_elements_t (model_two_body.hpp:40-62) */
    struct elements_t
    {
        double pomega = 0.0;        // argument of periapse
        double tau = 0.0;           // time of last periapse
        double cm_position_x = 0.0;
        double cm_position_y = 0.0;
        double cm_velocity_x = 0.0;
        double cm_velocity_y = 0.0;
        double separation = 1.0;
        double total_mass = 1.0;
        double mass_ratio = 1.0;
        double eccentricity = 0.0;
    };

    /** point_mass_t / two_body_state_t (model_two_body.hpp:66-82) */
    struct point_mass_t { double mass, x, y, vx, vy; };
    struct two_body_t { point_mass_t body1, body2; };

    M3B_HD elements_t elements_zero()
    {
        elements_t e;
        e.separation = e.total_mass = e.mass_ratio = e.eccentricity = 0.0;
        return e;
    }

    M3B_HD elements_t operator+(const elements_t& a, const elements_t& b)
    {
        return {a.pomega + b.pomega, a.tau + b.tau, a.cm_position_x + b.cm_position_x, a.cm_position_y + b.cm_position_y,
                a.cm_velocity_x + b.cm_velocity_x, a.cm_velocity_y + b.cm_velocity_y, a.separation + b.separation,
                a.total_mass + b.total_mass, a.mass_ratio + b.mass_ratio, a.eccentricity + b.eccentricity};
    }

    M3B_HD elements_t operator*(const elements_t& a, double s)
    {
        return {a.pomega * s, a.tau * s, a.cm_position_x * s, a.cm_position_y * s, a.cm_velocity_x * s,
                a.cm_velocity_y * s, a.separation * s, a.total_mass * s, a.mass_ratio * s, a.eccentricity * s};
    }

    M3B_HD double orbital_period(const elements_t& e)
    {
        const double two_pi = 6.283185307179586476925286766559;
        return two_pi / sqrt(e.total_mass / e.separation / e.separation / e.separation);
    }

    /** Positions and velocities in the orbit's own frame at time t past periapse. */
    M3B_HD two_body_t orbit_frame_state(const elements_t& el, double t)
    {
        double e = el.eccentricity, q = el.mass_ratio, a = el.separation, M = el.total_mass;
        double omega = a == 0.0 ? 0.0 : sqrt(M / a / a / a);
        double mu = q / (1.0 + q);
        double E = omega * t;       // eccentric anomaly; equals the mean anomaly for circular orbits

        if (e > 0.0)                // Newton-Raphson on Kepler's equation, |residual| <= 1e-10
        {
            double mean = E;
            double res = E - e * sin(E) - mean;
            while (fabs(res) > 1e-10)
            {
                E -= res / (1 - e * cos(E));
                res = E - e * sin(E) - mean;
            }
        }
        double sE = sin(E), cE = cos(E), k = sqrt(1 - e * e);
        two_body_t s;
        s.body1.mass = M * (1 - mu);
        s.body2.mass = M * mu;
        s.body1.x = -a * mu * (e - cE);
        s.body1.y = +a * mu * (0 + sE) * k;
        s.body2.x = -s.body1.x / q;
        s.body2.y = -s.body1.y / q;
        s.body1.vx = -a * mu * omega / (1 - e * cE) * sE;
        s.body1.vy = +a * mu * omega / (1 - e * cE) * cE * k;
        s.body2.vx = -s.body1.vx / q;
        s.body2.vy = -s.body1.vy / q;
        return s;
    }

    /** compute_two_body_state(full elements, t): rotate by pomega, translate by the CM state. */
    M3B_HD two_body_t two_body_state(const elements_t& el, double t)
    {
        while (t < el.tau) t += orbital_period(el);

        two_body_t L = orbit_frame_state(el, t - el.tau);
        double c = cos(-el.pomega), s = sin(-el.pomega);
        auto place = [&] (const point_mass_t& p)
        {
            point_mass_t r;
            r.mass = p.mass;
            r.x  = (+p.x * c + p.y * s) + el.cm_position_x;
            r.y  = (-p.x * s + p.y * c) + el.cm_position_y;
            r.vx = (+p.vx * c + p.vy * s) + el.cm_velocity_x;
            r.vy = (-p.vx * s + p.vy * c) + el.cm_velocity_y;
            return r;
        };
        return {place(L.body1), place(L.body2)};
    }

    /**
     * Inverse problem: orbital elements of a pair of point masses at time t.
     * Returns false (and leaves `out` untouched) if the pair is unbound, where the
     * reference throws std::invalid_argument (model_two_body.hpp:385-386).
     */
    M3B_HD bool orbital_elements(const two_body_t& s, double t, elements_t& out)
    {
        const point_mass_t& c1 = s.body1;
        const point_mass_t& c2 = s.body2;
        double M1 = c1.mass, M2 = c2.mass, M = M1 + M2, q = M2 / M1;
        double x_cm = (c1.x * c1.mass + c2.x * c2.mass) / M, y_cm = (c1.y * c1.mass + c2.y * c2.mass) / M;
        double vx_cm = (c1.vx * c1.mass + c2.vx * c2.mass) / M, vy_cm = (c1.vy * c1.mass + c2.vy * c2.mass) / M;
        double x1 = c1.x - x_cm, y1 = c1.y - y_cm, x2 = c2.x - x_cm, y2 = c2.y - y_cm;
        double r1 = sqrt(x1 * x1 + y1 * y1), r2 = sqrt(x2 * x2 + y2 * y2);
        double vx1 = c1.vx - vx_cm, vy1 = c1.vy - vy_cm, vx2 = c2.vx - vx_cm, vy2 = c2.vy - vy_cm;
        double vf1 = -vx1 * y1 / r1 + vy1 * x1 / r1;
        double vf2 = -vx2 * y2 / r2 + vy2 * x2 / r2;
        double v1 = sqrt(vx1 * vx1 + vy1 * vy1);
        double E1 = 0.5 * M1 * (vx1 * vx1 + vy1 * vy1), E2 = 0.5 * M2 * (vx2 * vx2 + vy2 * vy2);
        double L = M1 * r1 * vf1 + M2 * r2 * vf2;
        double E = E1 + E2 - M1 * M2 / (r1 + r2);

        if (E >= 0.0) return false;

        double a = -0.5 * M1 * M2 / E;
        double b = sqrt(-0.5 * L * L / E * (M1 + M2) / (M1 * M2));
        double e = sqrt(fmin(fmax(1.0 - b * b / a / a, 0.0), 1.0));
        double omega = sqrt(M / a / a / a);
        double a1 = a * q / (1.0 + q), b1 = b * q / (1.0 + q);
        double cn = e == 0.0 ? x1 / r1 : (1.0 - r1 / a1) / e;            // cos, sin of the phase angle
        double cf = a1 / r1 * (cn - e);                                 // ... and of the true anomaly
        double sn = e == 0.0 ? y1 / r1 : (vx1 * x1 + vy1 * y1) / (e * v1 * r1) * sqrt(1.0 - e * e * cn * cn);
        double sf = (b1 / r1) * sn;
        double cE = (e + cf) / (1.0 + e * cf);                          // eccentric anomaly
        double sE = sqrt(1.0 - e * e) * sf / (1.0 + e * cf);
        double mean_anomaly = atan2(sE, cE) - e * sE;
        double ax = +(cn - e) * x1 + sn * sqrt(1.0 - e * e) * y1;
        double ay = +(cn - e) * y1 - sn * sqrt(1.0 - e * e) * x1;

        out.pomega = atan2(ay, ax);
        out.tau = t - mean_anomaly / omega;
        out.cm_position_x = x_cm;
        out.cm_position_y = y_cm;
        out.cm_velocity_x = vx_cm;
        out.cm_velocity_y = vy_cm;
        out.separation = a;
        out.total_mass = M;
        out.mass_ratio = q;
        out.eccentricity = e;
        return true;
    }

    /** b - a with the angle / periapse-time differences wrapped to the nearest image. */
    M3B_HD elements_t elements_diff(const elements_t& a, const elements_t& b)
    {
        auto nearest = [] (double d, double period)
        {
            double lo = d - period, hi = d + period;
            if (fabs(d) < fmin(fabs(hi), fabs(lo))) return d;
            return fabs(hi) < fabs(lo) ? hi : lo;
        };
        const double two_pi = 6.283185307179586476925286766559;
        elements_t r;
        r.pomega = nearest(b.pomega - a.pomega, two_pi);
        r.tau = nearest(b.tau - a.tau, orbital_period(b));
        r.cm_position_x = b.cm_position_x - a.cm_position_x;
        r.cm_position_y = b.cm_position_y - a.cm_position_y;
        r.cm_velocity_x = b.cm_velocity_x - a.cm_velocity_x;
        r.cm_velocity_y = b.cm_velocity_y - a.cm_velocity_y;
        r.separation = b.separation - a.separation;
        r.total_mass = b.total_mass - a.total_mass;
        r.mass_ratio = b.mass_ratio - a.mass_ratio;
        r.eccentricity = b.eccentricity - a.eccentricity;
        return r;
    }

    /** Drift of the centre of mass over dt (diff_cm, model_two_body.hpp:522-529). */
    M3B_HD elements_t elements_cm_drift(const elements_t& a, double dt)
    {
        elements_t r = elements_zero();
        r.cm_position_x = a.cm_velocity_x * dt;
        r.cm_position_y = a.cm_velocity_y * dt;
        return r;
    }
}
