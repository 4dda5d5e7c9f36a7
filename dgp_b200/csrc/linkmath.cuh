// linkmath.cuh -- closed-form integrals of the linked-GP predictor shared by predict.cu and vecchia.cu.
// Restated from dgpsi/functions.py:453-494 (IJ_matern) and dgpsi/vecchia.py:915-988 (Jd, Jd0); the numpy
// twin that is pinned against the reference's golden vectors is oracle/dgp_oracle.py (Jd, Jd0, I_matern_dim).
#pragma once
#include "common.cuh"

namespace dgpb {

// Matern-2.5 correlation of a scalar distance a (already |.|) with length l  (functions.py:471)
__device__ __forceinline__ double matern_plain(double zX, double l) {
    double a = fabs(zX);
    return (1.0 + kSqrt5 * a / l + 5.0 * zX * zX / (3.0 * l * l)) * exp(-kSqrt5 * a / l);
}

// one-dimensional Matern I factor (functions.py:463-471)
__device__ __forceinline__ double I_matern_dim(double x, double zm, double zv, double l) {
    const double zX = zm - x;
    if (zv == 0.0) return matern_plain(zX, l);
    const double muA = zX - kSqrt5 * zv / l, muB = zX + kSqrt5 * zv / l;
    const double sv = sqrt(2.0 * zv);
    const double gg = sqrt(0.5 * zv / M_PI) / l;
    const double l2 = l * l;
    double tA = exp((5.0 * zv - 2.0 * kSqrt5 * l * zX) / (2.0 * l2)) *
                ((1.0 + kSqrt5 * muA / l + 5.0 * (muA * muA + zv) / (3.0 * l2)) * 0.5 * (1.0 + erf(muA / sv)) +
                 (kSqrt5 + (5.0 * muA) / (3.0 * l)) * gg * exp(-0.5 * muA * muA / zv));
    double tB = exp((5.0 * zv + 2.0 * kSqrt5 * l * zX) / (2.0 * l2)) *
                ((1.0 - kSqrt5 * muB / l + 5.0 * (muB * muB + zv) / (3.0 * l2)) * 0.5 * (1.0 + erf(-muB / sv)) +
                 (kSqrt5 - (5.0 * muB) / (3.0 * l)) * gg * exp(-0.5 * muB * muB / zv));
    return tA + tB;
}

// one-dimensional Matern J integral, X1 != X2 (vecchia.py:915-959) -- see oracle/dgp_oracle.py:Jd
static __device__ __noinline__ double Jd_dev(double X1, double X2, double zm, double zv, double l) {
    const double x1 = fmin(X1, X2), x2 = fmax(X1, X2);
    const double l2 = l * l, l3 = l2 * l, l4 = l2 * l2;
    const double i9l4 = 1.0 / (9.0 * l4);
    const double sv = sqrt(2.0 * zv);
    const double gg = sqrt(0.5 * zv / M_PI);
    const double x1s = x1 * x1, x2s = x2 * x2, x12 = x1 * x2, xs = x1 + x2;
    // region z > x2
    const double E30 = 1.0 + (25.0 * x1s * x2s - 3.0 * kSqrt5 * (3.0 * l3 + 5.0 * l * x12) * xs +
                              15.0 * l2 * (x1s + x2s + 3.0 * x12)) * i9l4;
    const double E31 = (18.0 * kSqrt5 * l3 + 15.0 * kSqrt5 * l * (x1s + x2s) - (75.0 * l2 + 50.0 * x12) * xs +
                        60.0 * kSqrt5 * l * x12) * i9l4;
    const double E32 = 5.0 * (5.0 * x1s + 5.0 * x2s + 15.0 * l2 - 9.0 * kSqrt5 * l * xs + 20.0 * x12) * i9l4;
    const double E33 = 10.0 * (3.0 * kSqrt5 * l - 5.0 * x1 - 5.0 * x2) * i9l4;
    const double E34 = 25.0 * i9l4;
    const double muC = zm - 2.0 * kSqrt5 * zv / l;
    const double muC2 = muC * muC;
    const double E3A31 = E30 + muC * E31 + (muC2 + zv) * E32 + (muC2 * muC + 3.0 * zv * muC) * E33 +
                         (muC2 * muC2 + 6.0 * zv * muC2 + 3.0 * zv * zv) * E34;
    const double E3A32 = E31 + (muC + x2) * E32 + (muC2 + 2.0 * zv + x2s + muC * x2) * E33 +
                         (muC2 * muC + x2s * x2 + x2 * muC2 + muC * x2s + 3.0 * zv * x2 + 5.0 * zv * muC) * E34;
    const double P1 = exp((10.0 * zv + kSqrt5 * l * (xs - 2.0 * zm)) / l2) *
                      (0.5 * E3A31 * (1.0 + erf((muC - x2) / sv)) + E3A32 * gg * exp(-0.5 * (x2 - muC) * (x2 - muC) / zv));
    // region x1 < z < x2
    const double E40 = 1.0 + (25.0 * x1s * x2s + 3.0 * kSqrt5 * (3.0 * l3 - 5.0 * l * x12) * (x2 - x1) +
                              15.0 * l2 * (x1s + x2s - 3.0 * x12)) * i9l4;
    const double E41 = 5.0 * (3.0 * kSqrt5 * l * (x2s - x1s) + 3.0 * l2 * xs - 10.0 * x12 * xs) * i9l4;
    const double E42 = 5.0 * (5.0 * x1s + 5.0 * x2s - 3.0 * l2 - 3.0 * kSqrt5 * l * (x2 - x1) + 20.0 * x12) * i9l4;
    const double E43 = -50.0 * (X1 + X2) * i9l4;
    const double E44 = 25.0 * i9l4;
    const double zm2 = zm * zm;
    const double E4A41 = E40 + zm * E41 + (zm2 + zv) * E42 + (zm2 * zm + 3.0 * zv * zm) * E43 +
                         (zm2 * zm2 + 6.0 * zv * zm2 + 3.0 * zv * zv) * E44;
    const double E4A42 = E41 + (zm + x1) * E42 + (zm2 + 2.0 * zv + x1s + zm * x1) * E43 +
                         (zm2 * zm + x1s * x1 + x1 * zm2 + zm * x1s + 3.0 * zv * x1 + 5.0 * zv * zm) * E44;
    const double E4A43 = E41 + (zm + x2) * E42 + (zm2 + 2.0 * zv + x2s + zm * x2) * E43 +
                         (zm2 * zm + x2s * x2 + x2 * zm2 + zm * x2s + 3.0 * zv * x2 + 5.0 * zv * zm) * E44;
    const double P2 = exp(-kSqrt5 * (x2 - x1) / l) *
                      (0.5 * E4A41 * (erf((x2 - zm) / sv) - erf((x1 - zm) / sv)) +
                       E4A42 * gg * exp(-0.5 * (x1 - zm) * (x1 - zm) / zv) -
                       E4A43 * gg * exp(-0.5 * (x2 - zm) * (x2 - zm) / zv));
    // region z < x1
    const double E50 = 1.0 + (25.0 * x1s * x2s + 3.0 * kSqrt5 * (3.0 * l3 + 5.0 * l * x12) * xs +
                              15.0 * l2 * (x1s + x2s + 3.0 * x12)) * i9l4;
    const double E51 = (18.0 * kSqrt5 * l3 + 15.0 * kSqrt5 * l * (x1s + x2s) + (75.0 * l2 + 50.0 * x12) * xs +
                        60.0 * kSqrt5 * l * x12) * i9l4;
    const double E52 = 5.0 * (5.0 * x1s + 5.0 * x2s + 15.0 * l2 + 9.0 * kSqrt5 * l * xs + 20.0 * x12) * i9l4;
    const double E53 = 10.0 * (3.0 * kSqrt5 * l + 5.0 * x1 + 5.0 * x2) * i9l4;
    const double E54 = 25.0 * i9l4;
    const double muD = zm + 2.0 * kSqrt5 * zv / l;
    const double muD2 = muD * muD;
    const double E5A51 = E50 - muD * E51 + (muD2 + zv) * E52 - (muD2 * muD + 3.0 * zv * muD) * E53 +
                         (muD2 * muD2 + 6.0 * zv * muD2 + 3.0 * zv * zv) * E54;
    const double E5A52 = E51 - (muD + x1) * E52 + (muD2 + 2.0 * zv + x1s + muD * x1) * E53 -
                         (muD2 * muD + x1s * x1 + x1 * muD2 + muD * x1s + 3.0 * zv * x1 + 5.0 * zv * muD) * E54;
    const double P3 = exp((10.0 * zv - kSqrt5 * l * (xs - 2.0 * zm)) / l2) *
                      (0.5 * E5A51 * (1.0 + erf((x1 - muD) / sv)) + E5A52 * gg * exp(-0.5 * (x1 - muD) * (x1 - muD) / zv));
    return P1 + P2 + P3;
}

// diagonal case X1 == X2 (vecchia.py:961-988)
static __device__ __noinline__ double Jd0_dev(double x1, double zm, double zv, double l) {
    const double l2 = l * l, l3 = l2 * l, l4 = l2 * l2;
    const double i9l4 = 1.0 / (9.0 * l4);
    const double sv = sqrt(2.0 * zv);
    const double gg = sqrt(0.5 * zv / M_PI);
    const double x1s = x1 * x1;
    const double E30 = 1.0 + (25.0 * x1s * x1s - 6.0 * kSqrt5 * (3.0 * l3 + 5.0 * l * x1s) * x1 + 75.0 * l2 * x1s) * i9l4;
    const double E31 = (18.0 * kSqrt5 * l3 + 90.0 * kSqrt5 * l * x1s - (150.0 * l2 + 100.0 * x1s) * x1) * i9l4;
    const double E32 = 5.0 * (30.0 * x1s + 15.0 * l2 - 18.0 * kSqrt5 * l * x1) * i9l4;
    const double E33 = 10.0 * (3.0 * kSqrt5 * l - 10.0 * x1) * i9l4;
    const double E34 = 25.0 * i9l4;
    const double muC = zm - 2.0 * kSqrt5 * zv / l;
    const double muC2 = muC * muC;
    const double E3A31 = E30 + muC * E31 + (muC2 + zv) * E32 + (muC2 * muC + 3.0 * zv * muC) * E33 +
                         (muC2 * muC2 + 6.0 * zv * muC2 + 3.0 * zv * zv) * E34;
    const double E3A32 = E31 + (muC + x1) * E32 + (muC2 + 2.0 * zv + x1s + muC * x1) * E33 +
                         (muC2 * muC + x1s * x1 + x1 * muC2 + muC * x1s + 3.0 * zv * x1 + 5.0 * zv * muC) * E34;
    const double P1 = exp((10.0 * zv + kSqrt5 * l * (2.0 * x1 - 2.0 * zm)) / l2) *
                      (0.5 * E3A31 * (1.0 + erf((muC - x1) / sv)) + E3A32 * gg * exp(-0.5 * (x1 - muC) * (x1 - muC) / zv));
    const double E50 = 1.0 + (25.0 * x1s * x1s + 6.0 * kSqrt5 * (3.0 * l3 + 5.0 * l * x1s) * x1 + 75.0 * l2 * x1s) * i9l4;
    const double E51 = (18.0 * kSqrt5 * l3 + 90.0 * kSqrt5 * l * x1s + (150.0 * l2 + 100.0 * x1s) * x1) * i9l4;
    const double E52 = 5.0 * (30.0 * x1s + 15.0 * l2 + 18.0 * kSqrt5 * l * x1) * i9l4;
    const double E53 = 10.0 * (3.0 * kSqrt5 * l + 10.0 * x1) * i9l4;
    const double E54 = 25.0 * i9l4;
    const double muD = zm + 2.0 * kSqrt5 * zv / l;
    const double muD2 = muD * muD;
    const double E5A51 = E50 - muD * E51 + (muD2 + zv) * E52 - (muD2 * muD + 3.0 * zv * muD) * E53 +
                         (muD2 * muD2 + 6.0 * zv * muD2 + 3.0 * zv * zv) * E54;
    const double E5A52 = E51 - (muD + x1) * E52 + (muD2 + 2.0 * zv + x1s + muD * x1) * E53 -
                         (muD2 * muD + x1s * x1 + x1 * muD2 + muD * x1s + 3.0 * zv * x1 + 5.0 * zv * muD) * E54;
    const double P3 = exp((10.0 * zv - kSqrt5 * l * (2.0 * x1 - 2.0 * zm)) / l2) *
                      (0.5 * E5A51 * (1.0 + erf((x1 - muD) / sv)) + E5A52 * gg * exp(-0.5 * (x1 - muD) * (x1 - muD) / zv));
    return P1 + P3;
}


}  // namespace dgpb
