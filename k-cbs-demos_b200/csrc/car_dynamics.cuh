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

// negation of both halves (ptxas folds it into the operand modifier of the consuming FFMA2)
__device__ __forceinline__ f32x2 neg2(f32x2 a)
{
    float lo, hi;
    upk(a, lo, hi);
    return pk(-lo, -hi);
}

// Per-pair constants of the packed integrator.
struct CarConsts2 {
    f32x2 td, ntd;          // +-tan((h/2) w): advances tan(phi) by half a sub-step
    f32x2 tf, ntf;          // +-tan(h w): by a whole sub-step
    f32x2 tstep, ntstep;    // +-tan(step * w): from one propagation step to the next
    f32x2 dU, dUh, dU6;     // increments of U = (h/2)(v / L) per sub-step / half sub-step, and a sixth of the former
    f32x2 ua, v0;           // acceleration and the parent's speed (v is an exact linear function of time)
    float kscale;           // (h/2) / L
    float cpos;             // L / 3: position offsets are accumulated in units of this (see car_step2)
};

__device__ __forceinline__ float tan_small(float d)   // |d| <= 0.1: the series is exact to 1e-10
{
    const float d2 = d * d;
    return d * fmaf(d2, fmaf(d2, fmaf(d2, 17.0f / 315.0f, 2.0f / 15.0f), 1.0f / 3.0f), 1.0f);
}

__device__ __forceinline__ CarConsts2 car_consts2(const DevWorld &w, const float ua[2], const float uw[2], const float v0[2])
{
    CarConsts2 k;
    const float h = w.h, hh = 0.5f * w.h;
    k.kscale = hh * w.inv_len;
    k.cpos = __fdividef(1.0f / 3.0f, w.inv_len);
    float td[2], tf[2], tp[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        td[e] = tan_small(hh * uw[e]);
        tf[e] = tan_small(h * uw[e]);
        tp[e] = tan_small(w.step * uw[e]);
    }
    k.td = pk(td[0], td[1]);
    k.ntd = pk(-td[0], -td[1]);
    k.tf = pk(tf[0], tf[1]);
    k.ntf = pk(-tf[0], -tf[1]);
    k.tstep = pk(tp[0], tp[1]);
    k.ntstep = pk(-tp[0], -tp[1]);
    const float du0 = k.kscale * h * ua[0], du1 = k.kscale * h * ua[1];
    k.dU = pk(du0, du1);
    k.dUh = pk(0.5f * du0, 0.5f * du1);
    k.dU6 = pk(du0 * (1.0f / 6.0f), du1 * (1.0f / 6.0f));
    k.ua = pk(ua[0], ua[1]);
    k.v0 = pk(v0[0], v0[1]);
    return k;
}

// Sub-steps (of the ten per propagation step) whose two reciprocals -- 1 / (1 - T td) and 1 / (1 - T tf) -- are taken
// from ONE reciprocal of their product (+3 packed multiplies, -2 MUFU per sub-step of two samples): the knob that balances
// the FMA pipe against the MUFU pipe.
#ifndef KCBS_RCP_SHARE_MASK
#define KCBS_RCP_SHARE_MASK 0x3ff
#endif

// State of two samples carried from one propagation step to the next.
struct CarPair2 {
    f32x2 T;           // tan(phi) at the start of the step
    f32x2 TH;          // heading at the start of the step (wrapped)
    f32x2 offx, offy;  // position offsets from the parents in units of CarConsts2::cpos
};

// One propagation step of two samples at once: n_sub classical RK4 sub-steps with the reference's stages (car_step is the
// scalar form).  U = (h/2) v / L is an exact linear sequence, K = U tan(phi) the (h/2)-scaled heading rate.  With K1, K2,
// K4 at the stages, the stage headings are theta1 + {0, K1, K2, 2 K2} and (h/6) v = (L/3) U, so in complex form
//   (h/6) sum_i w_i v_i e^{i theta_i} = (L/3) e^{i (theta1 + K2)} [U1 e^{-i K2} + 2 U2 e^{i (K1 - K2)} + 2 U2 + U4 e^{i K2}]
//     = (L/3) e^{i (theta1 + K2)} (A + i B),   A = U2 (6 - K2^2),   B = dU K2 + 2 U2 (K1 - K2)
// (U1 + U4 = 2 U2, U4 - U1 = dU exactly; |K2| <= 0.0175 and |K1 - K2| ~ 1e-4: the neglected terms are below 4e-8 A), and
// because B / A = dU tan(phi2) / 6 + (K1 - K2) / 3 is ~1e-4, A + i B = A e^{i B / A} to 1e-8: the four stages collapse to
//   A (cos, sin)(theta*),   theta* = theta1 + (K1 + 2 K2) / 3 + (dU / 6) tan(phi2)
// with ONE MUFU sin / cos per sample and sub-step (its 5e-7 error enters the position with weight h v and does not
// accumulate).  The heading advances by (K1 + 4 K2 + K4) / 3.  tan(phi) at the half and the full sub-step both follow
// from the start of the sub-step by the addition formula (independent reciprocals: no chain through the half step).
// 21 packed instructions + 2 FMUL + 8 MUFU (4 RCP, 2 SIN, 2 COS) per sub-step of two samples; an FFMA2 with three
// register operands occupies the FMA pipe for 3 cycles, every other packed instruction for 2 (tools/ubench/pipes.cu).
// `t_start` = time of the step's start (v there = v0 + ua t_start).  Returns the heading at the end of the step (not
// wrapped); p.T, p.offx, p.offy are advanced, p.TH is left to the caller (it wraps).
template <int N_SUB>
__device__ __forceinline__ f32x2 car_step2(const CarConsts2 &k, int n_sub, float t_start, CarPair2 &p)
{
    const f32x2 ONE = pk1(1.0f), NEG1 = pk1(-1.0f), TWO = pk1(2.0f), FOUR = pk1(4.0f), SIX = pk1(6.0f), THIRD = pk1(1.0f / 3.0f);
    f32x2 T = p.T;
    // tan(phi) at the end of the step straight from its start (one rounding-level error per step, off the sub-steps'
    // dependency chain)
    p.T = mul2(add2(T, k.tstep), rcp2(fma2(T, k.ntstep, ONE)));
    const f32x2 V = fma2(k.ua, pk1(t_start), k.v0);
    const f32x2 U1 = mul2(pk1(k.kscale), V);
    f32x2 U2 = add2(U1, k.dUh);
    f32x2 K1 = mul2(U1, T);
    f32x2 toff3 = pk1(0.0f);     // 3 x the heading offset from p.TH
    f32x2 offx = p.offx, offy = p.offy;

#pragma unroll kPropUnroll
    for (int n = 0; n < (N_SUB > 0 ? N_SUB : n_sub); ++n) {
        const f32x2 D2 = fma2(T, k.ntd, ONE), N2 = add2(T, k.td);
        const f32x2 D4 = fma2(T, k.ntf, ONE), N4 = add2(T, k.tf);
        f32x2 T2;
#ifdef KCBS_EXP_NORCP
        // timing experiment only (inaccurate): 1 / (1 - e) ~ 1 + e, no MUFU
        if (true) {
            T2 = mul2(N2, fma2(T, k.td, ONE));
            T = mul2(N4, fma2(T, k.tf, ONE));
        } else
#endif
        if (N_SUB > 0 && ((KCBS_RCP_SHARE_MASK >> (n % 10)) & 1)) {
            const f32x2 R = rcp2(mul2(D2, D4));
            T2 = mul2(mul2(N2, D4), R);
            T = mul2(mul2(N4, D2), R);
        } else {
            T2 = mul2(N2, rcp2(D2));
            T = mul2(N4, rcp2(D4));
        }
        const f32x2 U4 = add2(U2, k.dUh);
        const f32x2 K2 = mul2(U2, T2), K4 = mul2(U4, T);
        const f32x2 A = mul2(U2, fma2(mul2(K2, K2), NEG1, SIX));
        const f32x2 z = add2(toff3, K1);
        const f32x2 ang = fma2(fma2(K2, TWO, z), THIRD, fma2(k.dU6, T2, p.TH));   // theta*
        {
            float a0, a1, c0, c1, s0, s1;
            upk(ang, a0, a1);
            __sincosf(a0, &s0, &c0);
            __sincosf(a1, &s1, &c1);
            offx = fma2(pk(c0, c1), A, offx);
            offy = fma2(pk(s0, s1), A, offy);
        }
        toff3 = fma2(K2, FOUR, add2(z, K4));
        K1 = K4;
        U2 = add2(U2, k.dU);
    }
    p.offx = offx;
    p.offy = offy;
    return fma2(toff3, THIRD, p.TH);
}

}  // namespace kcbs
