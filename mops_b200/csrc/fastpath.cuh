// fastpath.cuh -- the straight-line form of one RK4 step on a hexagonal cell (the production hot path of k_advect).
//
// Why it exists (measured on B200, profiles/README.md "round 2"): the fp64 pipe issues one warp instruction per two cycles
// and a dependent fp64 instruction can issue 8 cycles after its producer; k_advect keeps only 3 warps per SM sub-partition
// resident (168 registers), so the pipe stays busy only while every warp has >= 2-3 independent fp64 instructions to issue.
// The per-instruction stall samples of the round-1 kernel show that it does inside the long unrolled blocks (2.3 cycles per
// fp64 instruction) and does not around every data-dependent branch: each range test of an exact div / sqrt sequence, each
// `if (status)` of the evaluation, each early-out of the rotation costs 10-16 cycles of branch latency and, worse, ends the
// basic block, so the independent chains on both sides of it cannot be interleaved by the scheduler.
//
// This file therefore restates the step WITHOUT data-dependent branches: every exact sequence runs unconditionally, every
// condition under which the reference would do something else (leave the cell, change layer, hit a degenerate operand, take
// nvcc's slow path of a division or root) is OR-ed into one `bad` word, and the caller tests it ONCE per step.  bad == 0
// means: every decision of the reference's step was the common one and every fast sequence was inside its exactness window,
// so the results are bit-identical to the generic code (engine.cuh / kernels.cuh), which the caller runs for the whole step
// when bad != 0 (about one step in 4000 on the bench workload).  Arithmetic is kept in the reference's association order;
// citations as in engine.cuh (VK = src/CPU/TBB/Kernel/MPASOVisualizerKernels.cpp, TK = .../TBBKernel.h).
#pragma once
#include "engine.cuh"

namespace mops {

#ifndef MOPS_FAST_UNROLL_SNAP
#define MOPS_FAST_UNROLL_SNAP 1
#endif
#ifndef MOPS_FAST_LOAD24
#define MOPS_FAST_LOAD24 1
#endif
#ifndef MOPS_FAST_LDG
#define MOPS_FAST_LDG 0 // 1 = explicit ld.global.nc for the (laundered) cell record instead of generic-space loads
#endif

// exact x / 6.0 of the RK4 combine; flags anything outside the window in which the correction step is exact
__device__ __forceinline__ double fast_div6(double a, unsigned& bad)
{
    const double x6 = 0x1.5555555555555p-3;
    const double q = a * x6;
    const double r = fma(-6.0, q, a);
    const bool z = is_zero(a);
    bad |= (unsigned)!(in_win_abs(a) || z);
    return z ? q : fma(x6, r, q);
}

// exact square root of a sum of squares (nvcc's fast-path sequence); flags operands outside its range (zero included)
__device__ __forceinline__ double fast_sqrt(double s, unsigned& bad)
{
    bad |= (unsigned)!(hi_raw(s) - 0x03500000u < 0x7ff00000u - 0x03500000u);
    return sq_fast(s);
}

// three exact quotients by one positive divisor (vector normalisation)
__device__ __forceinline__ void fast_div3(double a0, double a1, double a2, double b, double& q0, double& q1, double& q2, unsigned& bad)
{
    const double x = recip_refine(b);
    q0 = div_by(a0, b, x); q1 = div_by(a1, b, x); q2 = div_by(a2, b, x);
    const unsigned h0 = hi_abs(q0), h1 = hi_abs(q1), h2 = hi_abs(q2);
    const unsigned mn = min(min(h0, h1), h2), mx = max(max(h0, h1), h2);
    bad |= (unsigned)!(in_win_pos(b) && mn >= WIN_LO && mx < WIN_HI);
}

// advect_on_sphere (VK:729-738, TK:166-204) for the common case: |v| and |x| not tiny, |theta| < 2^-7, axis not
// degenerate.  x_rr = recip_refine(r_pos), hoisted by the caller (r_pos is the same for the three stage points of a step).
__device__ __forceinline__ d3 fast_rotate(const d3& pos, const d3& vel, double dt_local, double r_pos, double x_rr, unsigned& bad)
{
    d3 axis;
    axis.x = pos.y * vel.z - pos.z * vel.y;
    axis.y = pos.z * vel.x - pos.x * vel.z;
    axis.z = pos.x * vel.y - pos.y * vel.x;
    const double speed = fast_sqrt(vel.x * vel.x + vel.y * vel.y + vel.z * vel.z, bad);
    const double axis_len = fast_sqrt(axis.x * axis.x + axis.y * axis.y + axis.z * axis.z, bad);
    bad |= (unsigned)!(speed >= 1e-12);    // reference: speed < 1e-12 -> position unchanged (generic path)
    bad |= (unsigned)!(axis_len > 1e-12);  // reference: axis_len <= 1e-12 -> position unchanged
    const double num = speed * dt_local;
    const double theta = div_by(num, r_pos, x_rr);
    bad |= (unsigned)!(in_win_abs(theta)); // quotient inside the exactness window (r_pos is checked by the caller)
    bad |= (unsigned)!(fabs(theta) < 0.0078125);
    double sinTheta, cosTheta;
    {   // sincos_rot's small-angle branch
        const double x = theta * theta;
        double p = fma(x, 2.7557319223985893e-06, -1.9841269841269841e-04);
        p = fma(x, p, 8.3333333333333332e-03);
        p = fma(x, p, -1.6666666666666666e-01);
        sinTheta = fma(theta * x, p, theta);
        double q = fma(x, 2.4801587301587302e-05, -1.3888888888888889e-03);
        q = fma(x, q, 4.1666666666666664e-02);
        q = fma(x, q, -0.5);
        cosTheta = fma(x, q, 1.0);
    }
    d3 u;
    fast_div3(axis.x, axis.y, axis.z, axis_len, u.x, u.y, u.z, bad);
    d3 rotated;
    rotated.x = (cosTheta + u.x * u.x * (1.0 - cosTheta)) * pos.x +
        (u.x * u.y * (1.0 - cosTheta) - u.z * sinTheta) * pos.y +
        (u.x * u.z * (1.0 - cosTheta) + u.y * sinTheta) * pos.z;
    rotated.y = (u.y * u.x * (1.0 - cosTheta) + u.z * sinTheta) * pos.x +
        (cosTheta + u.y * u.y * (1.0 - cosTheta)) * pos.y +
        (u.y * u.z * (1.0 - cosTheta) - u.x * sinTheta) * pos.z;
    rotated.z = (u.z * u.x * (1.0 - cosTheta) - u.y * sinTheta) * pos.x +
        (u.z * u.y * (1.0 - cosTheta) + u.x * sinTheta) * pos.y +
        (cosTheta + u.z * u.z * (1.0 - cosTheta)) * pos.z;
    return rotated;
}

// one snapshot's share of calc_velocity_at on a hexagon with known layer `h` (= the previous evaluation's layer): checks
// that `h` is what the reference's search returns (engine.cuh, layer_search_stream / layer_search_path), forms t and
// gathers + blends levels h, h-1 (VK:1229-1286 / VK:828-870).
// NOW (how the vertical velocity of the call's snapshots is stored): 0 = present (or mixed: per-snapshot `w_is_z` decides),
// 1 = both snapshots without it, w slots hold +0.0: the vertical sums are exactly +0.0 and are not accumulated, 24-byte loads
// of (vx,vy,vz); 2 = both without it and the w slot of a record holds zTop of that (vertex, level): one 32-byte load per
// vertex and level serves velocity and layer probe (6 load instructions fewer per vertex than form 1, same bytes).
template <int M, bool PATH, int NOW>
__device__ __forceinline__ void fast_snapshot(const SnapView& s, const voff_t (&vo)[M], const double (&w)[M], int L, double depth, int hint,
                                              double& vx, double& vy, double& vz, double& vw, unsigned& bad)
{
    const double eps = 1e-8;
    const int h = min(max(hint, 1), L - 1);
    bad |= (unsigned)(h != hint);
    double top = 0.0, bot = 0.0;
    double dx = 0.0, dy = 0.0, dz = 0.0, dw = 0.0, ux = 0.0, uy = 0.0, uz = 0.0, uw = 0.0;
    if (NOW != 2) {
#pragma unroll
        for (int i = 0; i < M; ++i) { // VK:774-781 for levels h-1, h, vertex order
            const double* __restrict__ zq = s.ztop + (vo[i] + (voff_t)h);
            top += w[i] * zq[-1];
            bot += w[i] * zq[0];
        }
    }
    const bool wz = (NOW == 0) && (s.w_is_z != 0);
#pragma unroll
    for (int i = 0; i < M; ++i) { // TK:128-164 for the same two levels, vertex order
        const double4* __restrict__ q = s.velw + (vo[i] + (voff_t)h);
        double4 d, u;
        if (NOW == 1 && MOPS_FAST_LOAD24) { // the w component is +0.0 and unused: 24-byte loads
            ldg_d3of4(q, d.x, d.y, d.z);
            ldg_d3of4(q - 1, u.x, u.y, u.z);
            d.w = 0.0; u.w = 0.0;
        } else {
            d = ldg_d4(q);
            u = ldg_d4(q - 1);
        }
        if (NOW == 2) { // w slot = zTop of (vertex, level)
            top += w[i] * u.w;
            bot += w[i] * d.w;
        }
        dx += w[i] * d.x;
        dy += w[i] * d.y;
        dz += w[i] * d.z;
        if (NOW == 0) dw += w[i] * (wz ? 0.0 : d.w);
        ux += w[i] * u.x;
        uy += w[i] * u.y;
        uz += w[i] * u.z;
        if (NOW == 0) uw += w[i] * (wz ? 0.0 : u.w);
    }
    bool match = (depth <= top + eps) & (depth >= bot - eps);
    if (PATH) match = match & ((h == 1) | (depth < top - eps));              // first match of the linear scan (VK:1182-1218)
    else match = match & (depth > bot + eps) & (depth < top - eps);          // unique match of the bisection (VK:791-822)
    bad |= (unsigned)!match;
    const double mn = (top < depth) ? top : depth;
    const double x = (bot < mn) ? mn : bot;
    const double denom = top - bot;
    bad |= (unsigned)!(fabs(denom) >= 1e-12);
    const double num = x - bot;
    const double t = div_by(num, denom, recip_refine(denom));
    // exactness window of the quotient; num == 0 (depth at or below the layer bottom) gives an exact +0
    bad |= (unsigned)!(in_win_pos(denom) && (in_win_pos(t) || is_zero(num)));
    if (!PATH) { // zero-velocity rejects of the streamline (VK:845-847): anything near the threshold goes to the generic path
        bad |= (unsigned)!(dx * dx + dy * dy + dz * dz >= 2.0e-24);
        bad |= (unsigned)!(ux * ux + uy * uy + uz * uz >= 2.0e-24);
    }
    const double omt = 1.0 - t;
    vx = t * ux + omt * dx;
    vy = t * uy + omt * dy;
    vz = t * uz + omt * dz;
    vw = (NOW != 0) ? 0.0 : t * uw + omt * dw;
}

// calc_velocity_at on a hexagon (nv == M), layers given by the hints.  PATH: front/back blended with alpha (VK:1124-1327);
// else the streamline form (VK:740-872).
template <int M, bool PATH, int NOW>
__device__ __forceinline__ void fast_eval(const CellRec<M>* __restrict__ rec, const SnapView* __restrict__ sv, int L, const d3& p, double depth,
                                          double alpha, int hint_f, int hint_b, double& hx, double& hy, double& hz, double& vv, unsigned& bad)
{
    // IsInMesh (TK:40-53): any negative direction -> outside.  Sign bits are OR-ed (a -0.0 is sent to the generic path too).
    unsigned sgn = 0u;
    auto ld = [](const double* q) { return MOPS_FAST_LDG ? ldg_f64(q) : *q; };
#pragma unroll
    for (int k = 0; k < M; ++k) {
        const double direction = ld(&rec->nx[k]) * p.x + ld(&rec->ny[k]) * p.y + ld(&rec->nz[k]) * p.z;
        sgn |= hi_raw(direction);
    }
    bad |= sgn >> 31;
    double w[M];
    bool wok;
    hex_weights<M, MOPS_FAST_LDG != 0>(rec, p.x, p.y, p.z, w, wok); // a non-finite p fails its windows
    bad |= (unsigned)!wok;
    voff_t vo[M];
#pragma unroll
    for (int i = 0; i < M; ++i) vo[i] = (voff_t)(MOPS_FAST_LDG ? ldg_s32(&rec->vid[i]) : rec->vid[i]) * (voff_t)L;
    if (PATH) {
        // front and back through ONE rolled copy of the snapshot code (FAST_UNROLL_SNAP = 2 interleaves them: more
        // independent chains, twice the gathered records in flight)
        double fx = 0.0, fy = 0.0, fz = 0.0, fw = 0.0;
        const double oma = 1.0 - alpha;
        constexpr int US = MOPS_FAST_UNROLL_SNAP;
#pragma unroll US
        for (int k = 0; k < 2; ++k) {
            double bx, by, bz, bw;
            fast_snapshot<M, PATH, NOW>(sv[k], vo, w, L, depth, k ? hint_b : hint_f, bx, by, bz, bw, bad);
            if (k == 0) {
                fx = bx; fy = by; fz = bz; fw = bw;
            } else {
                hx = alpha * bx + oma * fx; // VK:1259
                hy = alpha * by + oma * fy;
                hz = alpha * bz + oma * fz;
                vv = (NOW != 0) ? 0.0 : alpha * bw + oma * fw; // VK:1286
            }
        }
    } else {
        fast_snapshot<M, PATH, NOW>(sv[0], vo, w, L, depth, hint_f, hx, hy, hz, vv, bad);
        bad |= (unsigned)!(hx * hx + hy * hy + hz * hz >= 2.0e-24); // VK:850-852
    }
}

struct FastStep {
    d3 new_pos;     // position after the step
    d3 hvel;        // (s1 + 2 s2 + 2 s3 + s4) / 6
    float depth_f;  // depth after the step (float round trip, R3)
};

// One whole RK4 step (VK:931-986 / VK:1399-1465) in the start-of-step cell `rec`; returns bad (0 = results valid).
template <int M, bool PATH, int NOW>
__device__ __forceinline__ unsigned fast_rk4_step(const CellRec<M>* __restrict__ rec, const SnapView* __restrict__ sv, int L, const d3& pos,
                                                  float depth_f, double alpha, double dalpha, int delta_t, int hint_f, int hint_b, FastStep& out)
{
    unsigned bad = 0u;
    const double dt = (double)delta_t;
    const double cur_depth = -1.0 * (double)depth_f;
    const double r = fast_sqrt(pos.x * pos.x + pos.y * pos.y + pos.z * pos.z, bad);
    bad |= (unsigned)!(r >= 1e-12) | (unsigned)!in_win_pos(r); // reference: rr < 1e-12 -> stage point = pos (generic path)
    const double x_rr = recip_refine(r);
    d3 acc = mk3(0.0, 0.0, 0.0), hprev = mk3(0.0, 0.0, 0.0);
    double vacc = 0.0;
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        d3 p = pos;
        double a_s = alpha;
        if (s > 0) {
            p = fast_rotate(pos, hprev, (s == 3) ? dt : dt * 0.5, r, x_rr, bad);
            if (PATH) a_s = clamp01(alpha + ((s == 3) ? dalpha : 0.5 * dalpha)); // VK:1410-1424
        }
        // the record is the same for the four stages: without this the compiler hoists all of its ~70 loads out of the stage
        // loop and parks them in local memory (measured: 450 B of spills); re-reading them per stage hits L1
        const CellRec<M>* __restrict__ rec_s = rec;
        asm volatile("" : "+l"(rec_s));
        double hx, hy, hz, vv;
        fast_eval<M, PATH, NOW>(rec_s, sv, L, p, cur_depth, a_s, hint_f, hint_b, hx, hy, hz, vv, bad);
        if (s == 0) {
            acc = mk3(hx, hy, hz);
            vacc = vv;
        } else {
            const double c = (s == 3) ? 1.0 : 2.0; // s1 + 2 s2 + 2 s3 + s4, left to right (VK:959-960)
            acc.x = acc.x + c * hx;
            acc.y = acc.y + c * hy;
            acc.z = acc.z + c * hz;
            if (NOW == 0) vacc = vacc + c * vv;
        }
        hprev = mk3(hx, hy, hz);
    }
    d3 hvel;
    hvel.x = fast_div6(acc.x, bad);
    hvel.y = fast_div6(acc.y, bad);
    hvel.z = fast_div6(acc.z, bad);
    const double vvel = (NOW != 0) ? 0.0 : fast_div6(vacc, bad);
    const double tx = pos.x + hvel.x * dt, ty = pos.y + hvel.y * dt, tz = pos.z + hvel.z * dt; // VK:962-964
    const double tl = fast_sqrt(tx * tx + ty * ty + tz * tz, bad);
    bad |= (unsigned)!(tl > 1e-12);
    double ux, uy, uz;
    fast_div3(tx, ty, tz, tl, ux, uy, uz, bad);
    d3 np = mk3(ux * r, uy * r, uz * r);
    // depth / radius update with the float round trip (VK:977-986, R3, R4)
    const double old_depth = (double)depth_f;
    double new_depth = old_depth - vvel * (double)delta_t;
    new_depth = (0.0 < new_depth) ? new_depth : 0.0;
    const double r_sum = r + vvel * (double)delta_t;
    const double r_new = (1.0 < r_sum) ? r_sum : 1.0;
    out.depth_f = (float)new_depth;
    const double nlen = fast_sqrt(np.x * np.x + np.y * np.y + np.z * np.z, bad);
    bad |= (unsigned)!(nlen > 1e-12);
    fast_div3(np.x, np.y, np.z, nlen, ux, uy, uz, bad);
    out.new_pos = mk3(ux * r_new, uy * r_new, uz * r_new);
    out.hvel = hvel;
    return bad;
}

} // namespace mops
