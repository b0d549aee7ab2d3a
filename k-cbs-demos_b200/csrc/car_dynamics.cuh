// car_dynamics.cuh -- one 0.1 s propagation step of the 2nd-order car (SecondOrderCarODE,
// StatePropagatorDatabase.h:44-64) as classical RK4 with sub-step h (odeint runge_kutta4 inside ODEBasicSolver),
// shared by the single-car expansion kernel (propagate.cu) and the merged-pair kernels (pair.cu).
// See the header of propagate.cu for the derivation of the stage sharing and the rotation recurrences.
#pragma once

#include "kcbs_device.cuh"

namespace kcbs {

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Per-sample constants of the integrator (the control is constant over a segment).
struct CarConsts {
    float td;        // tan((h/2) w): half-sub-step increment of phi
    float dUh, dU;   // increments of U = (h/2)(v / L) per half / full sub-step
    float dW2, dW;   // increments of W = (h/6) v
    float kscale;    // (h/2) / L
    float h6;        // h / 6
};

__device__ __forceinline__ CarConsts car_consts(const DevWorld &w, float ua, float uw)
{
    CarConsts k;
    const float h = w.h, hh = 0.5f * w.h;
    k.h6 = w.h * (1.0f / 6.0f);
    k.kscale = hh * w.inv_len;
    const float dl = hh * uw, dl2 = dl * dl;
    k.td = dl * fmaf(dl2, fmaf(dl2, 2.0f / 15.0f, 1.0f / 3.0f), 1.0f);
    k.dUh = k.kscale * hh * ua;
    k.dU = k.kscale * h * ua;
    k.dW2 = 2.0f * k.h6 * hh * ua;
    k.dW = k.h6 * h * ua;
    return k;
}

// One propagation step (n_sub RK4 sub-steps) from speed vs, steering angle ps, heading ts: adds the position
// increment to (offx, offy) and returns the heading increment (before the wrap).
__device__ __forceinline__ float car_step(const CarConsts &k, int n_sub, float vs, float ps, float ts, float &offx,
                                          float &offy)
{
    float sp, cp, s1, c1;
    sincosf(ps, &sp, &cp);
    __sincosf(ts, &s1, &c1);
    float T = sp * rcp_approx(cp);
    float U1 = k.kscale * vs, W1 = k.h6 * vs;
    float K1 = U1 * T;   // (h/2) k_theta at stage 1 of sub-step 0
    float toff = 0.0f;   // heading offset from ts within this step

#pragma unroll 2
    for (int n = 0; n < n_sub; ++n) {
        const float U2 = U1 + k.dUh, U4 = U1 + k.dU;
        const float W2 = fmaf(2.0f, W1, k.dW2), W4 = W1 + k.dW;
        // tan(phi) at the half and the full sub-step
        const float T2 = (T + k.td) * rcp_approx(fmaf(-T, k.td, 1.0f));
        T = (T2 + k.td) * rcp_approx(fmaf(-T2, k.td, 1.0f));
        const float K2 = U2 * T2, K4 = U4 * T;

        // stage 2 heading = stage 1 + K1 (2nd-order rotation)
        const float h2 = -0.5f * K1;
        const float c2 = fmaf(fmaf(c1, h2, -s1), K1, c1);
        const float s2 = fmaf(fmaf(s1, h2, c1), K1, s1);
        // stage 3 heading = stage 2 + (K2 - K1): only cos2 + cos3 and sin2 + sin3 are needed
        const float e3 = K2 - K1;
        const float c23 = fmaf(-s2, e3, c2 + c2);
        const float s23 = fmaf(c2, e3, s2 + s2);
        // stage 4 heading = stage 1 + 2 K2 (k3 == k2) = stage 2 + (2 K2 - K1)
        const float e4 = fmaf(2.0f, K2, -K1), h4 = -0.5f * e4;
        const float c4 = fmaf(fmaf(c2, h4, -s2), e4, c2);
        const float s4 = fmaf(fmaf(s2, h4, c2), e4, s2);

        offx = fmaf(W1, c1, fmaf(W2, c23, fmaf(W4, c4, offx)));
        offy = fmaf(W1, s1, fmaf(W2, s23, fmaf(W4, s4, offy)));

        // heading: theta += h/6 (k1 + 4 k2 + k4) = (K1 + 4 K2 + K4) / 3; next stage 1 = stage 4 + (that - 2 K2)
        const float dth = (1.0f / 3.0f) * fmaf(4.0f, K2, K1 + K4);
        toff += dth;
        const float e1 = fmaf(-2.0f, K2, dth);
        c1 = fmaf(-s4, e1, c4);
        s1 = fmaf(c4, e1, s4);
        K1 = K4;
        U1 = U4;
        W1 = W4;
    }
    return toff;
}

}  // namespace kcbs
