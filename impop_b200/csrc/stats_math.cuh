// Scalar fp64 statistics shared by the window and auxiliary kernels.  Every operation is an
// explicit round-to-nearest intrinsic, in the reference's order (citations to /root/reference/scripts).
#pragma once
#include "common.cuh"

namespace impop {

// Harmonic tables a1(n), a2(n) exactly as Python >= 3.12 builtin sum() (Neumaier) forms them in
// tj_d.py:41-45.  Sequential by nature; run once per context.
__device__ __forceinline__ void neumaier_add(double &total, double &comp, double x) {
    double t = __dadd_rn(total, x);
    if (fabs(total) >= fabs(x)) comp = __dadd_rn(comp, __dadd_rn(__dadd_rn(total, -t), x));
    else comp = __dadd_rn(comp, __dadd_rn(__dadd_rn(x, -t), total));
    total = t;
}
__device__ __forceinline__ double neumaier_value(double total, double comp) {
    return (comp != 0.0 && isfinite(comp)) ? __dadd_rn(total, comp) : total;
}

// tj_d.py:47-69 given a1, a2.  parts (nullable): a1 a2 b1 b2 c1 c2 e1 e2 numerator denominator.
__device__ inline double tajima_from_harmonics(double dn, double S, double pi, double a1, double a2, double *parts) {
    double b1 = __ddiv_rn(__dadd_rn(dn, 1.0), __dmul_rn(3.0, __dadd_rn(dn, -1.0)));
    double b2 = __ddiv_rn(__dmul_rn(2.0, __dadd_rn(__dadd_rn(__dmul_rn(dn, dn), dn), 3.0)),
                          __dmul_rn(__dmul_rn(9.0, dn), __dadd_rn(dn, -1.0)));
    double c1 = __dadd_rn(b1, -__ddiv_rn(1.0, a1));
    double c2 = __dadd_rn(__dadd_rn(b2, -__ddiv_rn(__dadd_rn(dn, 2.0), __dmul_rn(a1, dn))),
                          __ddiv_rn(a2, __dmul_rn(a1, a1)));
    double e1 = __ddiv_rn(c1, a1);
    double e2 = __ddiv_rn(c2, __dadd_rn(__dmul_rn(a1, a1), a2));
    double num = __dadd_rn(pi, -__ddiv_rn(S, a1));
    double den = (S > 0.0)
                     ? __dsqrt_rn(__dadd_rn(__dmul_rn(e1, S), __dmul_rn(__dmul_rn(e2, S), __dadd_rn(S, -1.0))))
                     : __longlong_as_double(0x7ff8000000000000ll);
    double d = (den != 0.0 && den == den) ? __ddiv_rn(num, den) : __longlong_as_double(0x7ff8000000000000ll);
    if (parts) {
        parts[0] = a1; parts[1] = a2; parts[2] = b1; parts[3] = b2; parts[4] = c1;
        parts[5] = c2; parts[6] = e1; parts[7] = e2; parts[8] = num; parts[9] = den;
    }
    return d;
}

// Derived statistics from raw sums; mirrors oracle_finalize() (oracle/csrc/oracle_impop.c).
__device__ inline void finalize_row(const double sums[4], const int64_t cnt[IMPOP_NCOUNTS], int64_t L, double seg,
                             const double2 *harm, int32_t harm_n, double *st) {
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const int64_t nS = cnt[0];
    double pi = 0.0;
    if (nS >= 2 && cnt[3] > 0) {
        double dn = (double)nS;
        double f = __ddiv_rn(1.0, dn);                                                     // pica2.py:137-138
        pi = __dmul_rn(__ddiv_rn(dn, __dadd_rn(dn, -1.0)),
                       __dmul_rn(2.0, __dmul_rn(__dmul_rn(sums[0], f), f)));               // pica2.py:154
    }
    double pps = (L > 0) ? __ddiv_rn(pi, (double)L) : nan;                                 // pica2.py:163-164
    double pi_a = cnt[4] > 0 ? __ddiv_rn(sums[1], (double)cnt[4]) : 0.0;                   // h-fst.py:168-171
    double pi_b = cnt[5] > 0 ? __ddiv_rn(sums[2], (double)cnt[5]) : 0.0;
    double pi_xy = __dmul_rn(0.5, __dadd_rn(pi_a, pi_b));                                  // h-fst.py:203
    double dxy = cnt[6] > 0 ? __ddiv_rn(sums[3], (double)cnt[6]) : 0.0;
    double fst = dxy > 0.0 ? __ddiv_rn(__dadd_rn(dxy, -pi_xy), dxy) : 0.0;                 // h-fst.py:214-222
    double da = __dadd_rn(dxy, -pi_xy);
    if (L > 0) {                                                                           // h-fst.py:225-240
        double dl = (double)L;
        pi_a = __ddiv_rn(pi_a, dl); pi_b = __ddiv_rn(pi_b, dl); pi_xy = __ddiv_rn(pi_xy, dl);
        dxy = __ddiv_rn(dxy, dl); da = __ddiv_rn(da, dl);
    }
    double parts[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) parts[k] = nan;
    double d = nan, d_raw = nan;
    if (nS >= 2) {
        double a1, a2;
        if (nS <= harm_n) { a1 = harm[nS].x; a2 = harm[nS].y; }
        else {   // beyond the table: form the sums directly
            double t1 = 0.0, c1 = 0.0, t2 = 0.0, c2 = 0.0;
            for (int64_t i = 1; i < nS; ++i) {
                double di = (double)i;
                neumaier_add(t1, c1, __ddiv_rn(1.0, di));
                neumaier_add(t2, c2, __ddiv_rn(1.0, __dmul_rn(di, di)));
            }
            a1 = neumaier_value(t1, c1); a2 = neumaier_value(t2, c2);
        }
        d_raw = tajima_from_harmonics((double)nS, seg, pi, a1, a2, parts);
        d = (L > 0) ? tajima_from_harmonics((double)nS, seg, pps, a1, a2, nullptr) : d_raw;  // run_tajd.sh:166-180
    }
    st[0] = pi; st[1] = pps; st[2] = pi_a; st[3] = pi_b; st[4] = pi_xy; st[5] = dxy; st[6] = da; st[7] = fst;
    st[8] = seg; st[9] = d; st[10] = parts[0]; st[11] = parts[6]; st[12] = parts[7]; st[13] = (double)nS;
    st[14] = sums[0]; st[15] = sums[1]; st[16] = sums[2]; st[17] = sums[3]; st[18] = d_raw; st[19] = 0.0;
}


}  // namespace impop
