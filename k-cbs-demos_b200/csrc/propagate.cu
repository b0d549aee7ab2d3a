// propagate.cu -- K1: fused multi-step RK4 expansion kernel (one thread per control sample).
//
// Replaces, per sample, oc::SpaceInformation::propagateWhileValid driving
//   SecondOrderCarODE                      StatePropagatorDatabase.h:44-64
//   odeint runge_kutta4, dt = 0.01         (ODEBasicSolver, call site demo_*.cpp:111-112)
//   SecondOrderCarODEPostIntegration       StatePropagatorDatabase.h:67-74 (theta wrap per step)
//   homogeneous2ndOrderCarSystemSVC::isValid  StateValidityCheckerDatabase.h:54-75
//
// The kernel is classical RK4 with the same stages as the reference, but it uses what the ODE's
// structure gives for free:
//   * vdot = a and phidot = w are constant, so the v/phi stages are v + a*c_i*h exactly and RK4
//     reproduces v(t), phi(t) exactly; stages 2 and 3 share v, phi and therefore tan(phi) and
//     k_theta; stage 4 of a sub-step equals stage 1 of the next one.
//   * phi advances by the constant (h/2)*w per half sub-step, so (sin phi, cos phi) are advanced with
//     a fixed rotation, re-anchored with an accurate sincosf at every propagation step.
//   * positions and the heading are accumulated as small offsets from the segment / step start so
//     FP32 rounding happens once per written state, not once per sub-step.
#include "kcbs_device.cuh"
#include "kernels.h"

namespace kcbs {

template <bool HAS_OBS>
__global__ void __launch_bounds__(256)
k_propagate(const __grid_constant__ DevWorld w, const __grid_constant__ PropArgs a)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;

    const long long pi = a.parent_idx ? (long long)a.parent_idx[i] : (a.n_parents == 1 ? 0 : i);
    const float x0 = a.parents[0 * a.n_parents + pi];
    const float y0 = a.parents[1 * a.n_parents + pi];
    const float v0 = a.parents[2 * a.n_parents + pi];
    const float phi0 = a.parents[3 * a.n_parents + pi];
    const float th0 = a.parents[4 * a.n_parents + pi];
    const float ua = a.controls[i];
    const float uw = a.controls[a.n + i];
    int st = a.steps ? a.steps[i] : a.max_steps;
    st = min(max(st, 0), a.max_steps);

    const float h = w.h, hh = 0.5f * w.h, h6 = w.h * (1.0f / 6.0f);
    const float hl = w.rob_hl[a.robot], hw = w.rob_hw[a.robot], rad = w.rob_rad[a.robot];

    // rotation by delta = (h/2) * w for the (sin phi, cos phi) recurrence: sin(delta) and 1 - cos(delta)
    const float dl = hh * uw, dl2 = dl * dl;
    const float sd = dl * fmaf(dl2, fmaf(dl2, 1.0f / 120.0f, -1.0f / 6.0f), 1.0f);
    const float ed = 0.5f * dl2 * fmaf(dl2, fmaf(dl2, 1.0f / 360.0f, -1.0f / 12.0f), 1.0f);

    float offx = 0.0f, offy = 0.0f;                  // position offset from the parent
    float xs = x0, ys = y0, vs = v0, ps = phi0, ts = th0;  // state at the start of the current step
    int nv = 0;

    for (int j = 0; j < st; ++j) {
        float sp, cp;
        sincosf(ps, &sp, &cp);
        float k1 = vs * w.inv_len * __fdividef(sp, cp);  // k_theta at stage 1 of sub-step 0
        float toff = 0.0f;                               // heading offset from ts within this step

        for (int n = 0; n < w.n_sub; ++n) {
            const float t0 = (float)n * h;
            const float v1 = fmaf(ua, t0, vs);
            const float v2 = fmaf(ua, t0 + hh, vs);
            const float v4 = fmaf(ua, t0 + h, vs);
            // phi -> phi + delta -> phi + 2 delta
            float s2 = fmaf(cp, sd, fmaf(-sp, ed, sp));
            float c2 = fmaf(-sp, sd, fmaf(-cp, ed, cp));
            const float k2 = v2 * w.inv_len * __fdividef(s2, c2);
            sp = fmaf(c2, sd, fmaf(-s2, ed, s2));
            cp = fmaf(-s2, sd, fmaf(-c2, ed, c2));
            const float k4 = v4 * w.inv_len * __fdividef(sp, cp);

            float sa, ca, sb, cb, sc, cc, se, ce;
            __sincosf(ts + toff, &sa, &ca);                 // stage 1
            __sincosf(ts + fmaf(hh, k1, toff), &sb, &cb);   // stage 2
            __sincosf(ts + fmaf(hh, k2, toff), &sc, &cc);   // stage 3
            __sincosf(ts + fmaf(h, k2, toff), &se, &ce);    // stage 4 (k3 == k2)

            offx = fmaf(h6, fmaf(v1, ca, fmaf(2.0f * v2, cb + cc, v4 * ce)), offx);
            offy = fmaf(h6, fmaf(v1, sa, fmaf(2.0f * v2, sb + sc, v4 * se)), offy);
            toff = fmaf(h6, fmaf(4.0f, k2, k1 + k4), toff);
            k1 = k4;
        }

        const float t1 = (float)(j + 1) * w.step;
        const float xn = x0 + offx, yn = y0 + offy;
        const float vn = fmaf(ua, t1, v0), pn = fmaf(uw, t1, phi0);
        const float tn = wrap_angle(ts + toff);

        bool ok = in_bounds(w, xn, yn, vn, pn, tn);
        if (HAS_OBS) {
            if (ok) ok = obstacles_clear(w, xn, yn, tn, hl, hw, rad);
        }
        if (!ok) break;

        xs = xn; ys = yn; vs = vn; ps = pn; ts = tn;
        nv = j + 1;
        if (a.seg) {
            const long long plane = (long long)a.max_steps * a.n;
            float *o = a.seg + (long long)j * a.n + i;
            __stcs(o, xn);
            __stcs(o + plane, yn);
            __stcs(o + 2 * plane, vn);
            __stcs(o + 3 * plane, pn);
            __stcs(o + 4 * plane, tn);
        }
    }

    if (a.valid_steps) a.valid_steps[i] = nv;
    if (a.fin) {
        a.fin[i] = xs;
        a.fin[a.n + i] = ys;
        a.fin[2 * a.n + i] = vs;
        a.fin[3 * a.n + i] = ps;
        a.fin[4 * a.n + i] = ts;
    }
}

cudaError_t launch_propagate(const DevWorld &w, const PropArgs &a, cudaStream_t stream)
{
    if (a.n <= 0) return cudaSuccess;
    const int block = 256;
    const long long grid = (a.n + block - 1) / block;
    if (w.n_obs > 0)
        k_propagate<true><<<(unsigned)grid, block, 0, stream>>>(w, a);
    else
        k_propagate<false><<<(unsigned)grid, block, 0, stream>>>(w, a);
    return cudaGetLastError();
}

}  // namespace kcbs
