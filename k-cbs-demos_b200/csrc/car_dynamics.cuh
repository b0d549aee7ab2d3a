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

#ifndef KCBS_PROP_UNROLL
#define KCBS_PROP_UNROLL 10
#endif
constexpr int kPropUnroll = KCBS_PROP_UNROLL;  // sub-steps per unrolled loop body of car_step2

// ---- two samples per thread in packed FP32 pairs (sm_100 fma/mul/add/sub .f32x2: one issue slot per two FMAs) ---------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 pk1(float v) { return pk(v, v); }
__device__ __forceinline__ void upk(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 rcp2(f32x2 a)
{
    float lo, hi;
    upk(a, lo, hi);
    return pk(rcp_approx(lo), rcp_approx(hi));
}

// Per-pair constants of the packed integrator.
struct CarConsts2 {
    f32x2 td, ntd;   // +-tan((h/2) w)
    f32x2 tf, ntf;   // +-tan(h w)
    f32x2 dU;        // increment of U = (h/2)(v / L) per sub-step
    f32x2 dW, dW2;   // increments of W = (h/6) v and of W2 = 2 W(half step) per sub-step
    f32x2 dWs;       // increment of Ws = W1 + 2 W2 + W4
    float tstep[2];  // tan(step * w): advances tan(phi) from one propagation step to the next
    float kscale, h6, hh_ua_k[2], hh_ua_w[2];
};

__device__ __forceinline__ CarConsts2 car_consts2(const DevWorld &w, const float ua[2], const float uw[2])
{
    CarConsts2 k;
    const float h = w.h, hh = 0.5f * w.h;
    k.h6 = w.h * (1.0f / 6.0f);
    k.kscale = hh * w.inv_len;
    float td[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const float dl = hh * uw[e], dl2 = dl * dl;
        td[e] = dl * fmaf(dl2, fmaf(dl2, 2.0f / 15.0f, 1.0f / 3.0f), 1.0f);
        const float ds = w.step * uw[e], ds2 = ds * ds;  // |ds| <= 0.1: the series is exact to 1e-10
        k.tstep[e] = ds * fmaf(ds2, fmaf(ds2, fmaf(ds2, 17.0f / 315.0f, 2.0f / 15.0f), 1.0f / 3.0f), 1.0f);
        k.hh_ua_k[e] = k.kscale * hh * ua[e];   // U at the half sub-step = U1 + this
        k.hh_ua_w[e] = 2.0f * k.h6 * hh * ua[e];  // W2 = 2 W1 + this
    }
    k.td = pk(td[0], td[1]);
    k.ntd = pk(-td[0], -td[1]);
    {
        float tf[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float df = h * uw[e], df2 = df * df;
            tf[e] = df * fmaf(df2, fmaf(df2, 2.0f / 15.0f, 1.0f / 3.0f), 1.0f);
        }
        k.tf = pk(tf[0], tf[1]);
        k.ntf = pk(-tf[0], -tf[1]);
    }
    k.dU = pk(k.kscale * h * ua[0], k.kscale * h * ua[1]);
    k.dW = pk(k.h6 * h * ua[0], k.h6 * h * ua[1]);
    k.dW2 = pk(2.0f * k.h6 * h * ua[0], 2.0f * k.h6 * h * ua[1]);
    k.dWs = pk(6.0f * k.h6 * h * ua[0], 6.0f * k.h6 * h * ua[1]);
    return k;
}

// One propagation step of two samples at once (same RK4 stages as car_step).  With K1, K2, K4 the (h/2)-scaled heading
// rates at the stages, the stage headings are theta1 + {0, K1, K2, 2 K2}, so
//   sum_i w_i cos(theta_i) = c1 A - s1 B,  sum_i w_i sin(theta_i) = s1 A + c1 B,
//   A = W1 + W2 (cos K1 + cos K2) + W4 cos 2K2,  B = W2 (sin K1 + sin K2) + W4 sin 2K2   (|K| <= 0.035:
//   cos d = 1 - d^2/2, sin d = d to the accuracy of car_step's second-order rotations),
// and stage 1 of the next sub-step takes (c1, s1) of theta + the accumulated dtheta = (K1 + 4 K2 + K4) / 3 from MUFU.
// 34 packed instructions + 8 MUFU (4 RCP, 2 SIN, 2 COS) per sub-step of two samples.
template <int N_SUB>
__device__ __forceinline__ f32x2 car_step2(const CarConsts2 &k, int n_sub, const float vs[2], const float T0[2],
                                           const float ts[2], f32x2 &offx, f32x2 &offy)
{
    float s1[2], c1[2], U0[2], W0[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        __sincosf(ts[e], &s1[e], &c1[e]);
        U0[e] = k.kscale * vs[e];
        W0[e] = k.h6 * vs[e];
    }
    const f32x2 ONE = pk1(1.0f), NHALF = pk1(-0.5f), FOUR = pk1(4.0f), THIRD = pk1(1.0f / 3.0f);
    f32x2 T = pk(T0[0], T0[1]);
    f32x2 U1 = pk(U0[0], U0[1]), U2 = pk(U0[0] + k.hh_ua_k[0], U0[1] + k.hh_ua_k[1]);
    f32x2 W1 = pk(W0[0], W0[1]), W2 = pk(fmaf(2.0f, W0[0], k.hh_ua_w[0]), fmaf(2.0f, W0[1], k.hh_ua_w[1]));
    f32x2 C = pk(c1[0], c1[1]), S = pk(s1[0], s1[1]);
    f32x2 K1 = mul2(U1, T);
    f32x2 Ws = add2(add2(W1, add2(W1, k.dW)), add2(W2, W2));
    f32x2 toff = pk1(0.0f);
    const f32x2 TS = pk(ts[0], ts[1]);

#pragma unroll kPropUnroll
    for (int n = 0; n < (N_SUB > 0 ? N_SUB : n_sub); ++n) {
        // tan(phi) at the half and the full sub-step
        const f32x2 T2 = mul2(add2(T, k.td), rcp2(fma2(T, k.ntd, ONE)));
        // tan(phi) at the full sub-step straight from the start of the sub-step (addition formula with tan(h w)): the two
        // reciprocals of a sub-step are independent, which halves the loop-carried dependency chain (4 % faster and a
        // slightly smaller error than chaining T2 -> T)
        T = mul2(add2(T, k.tf), rcp2(fma2(T, k.ntf, ONE)));
        const f32x2 U4 = add2(U1, k.dU), W4 = add2(W1, k.dW);
        const f32x2 K2 = mul2(U2, T2), K4 = mul2(U4, T);
        const f32x2 K22 = add2(K2, K2);
        // A = (W1 + 2 W2 + W4) - (W2 (K1^2 + K2^2) + W4 (2 K2)^2) / 2
        const f32x2 sq = fma2(K2, K2, mul2(K1, K1));
        const f32x2 m = fma2(W4, mul2(K22, K22), mul2(W2, sq));
        const f32x2 A = fma2(NHALF, m, Ws);
        const f32x2 B = fma2(W4, K22, mul2(W2, add2(K1, K2)));
        offx = sub2(fma2(C, A, offx), mul2(S, B));
        offy = fma2(C, B, fma2(S, A, offy));
        // heading
        const f32x2 dth = mul2(fma2(FOUR, K2, add2(K1, K4)), THIRD);
        toff = add2(toff, dth);
        {
            // stage-1 heading of the next sub-step straight from the angle: the MUFU pipe has room (4 RCP per sub-step
            // of two samples), the FMA pipe is the one that binds, and a rotation of (C, S) costs it ten instructions;
            // the MUFU error (5e-7) enters the position with weight W = (h/6) v and does not accumulate
            float a0, a1, c0, c1, s0, s1;
            upk(add2(TS, toff), a0, a1);
            __sincosf(a0, &s0, &c0);
            __sincosf(a1, &s1, &c1);
            C = pk(c0, c1);
            S = pk(s0, s1);
        }
        K1 = K4;
        U1 = U4;
        U2 = add2(U2, k.dU);
        W1 = W4;
        W2 = add2(W2, k.dW2);
        Ws = add2(Ws, k.dWs);
    }
    return toff;
}

}  // namespace kcbs
