// dpb_host.h -- host-side derivation of the equation / scheme constants from dpb_config.
// Constants are formed in double exactly as the reference's Python expressions are
// (equation.py:48-49,75,80,85-86,151,246,272,282,289,299), then narrowed to the compute type.
#pragma once
#include <math.h>
#include "../../include/deeppde_b200.h"
#include "dpb_eqn.h"

namespace dpb {

inline void fill_eqn(const dpb_config& c, int N, double T, EqnD& e) {
    e.eqn = c.eqn; e.d = c.dim; e.m = c.control_dim; e.scheme = c.scheme; e.td = c.td_type;
    e.R = c.R; e.R2 = c.R * c.R; e.gamma = c.discount;
    e.p = c.p; e.q = c.q; e.beta = c.beta; e.a = c.a; e.eps = c.epsilon; e.a2 = c.a2; e.a3 = c.a3;
    e.k = 0; e.cu = 0; e.wconst = 0; e.ZR = 0; e.C0 = 0;
    e.lv_num = e.lv_den = e.lv_un = e.lv_ud = e.lv_ue = e.lv_gk = 0;
    const double sigU = sqrt(2.0);                                        // equation.py:75 (sigma_Up)
    e.sig = sqrt(2.0);                                                    // equation.py:170,230,268,305
    const double d = (double)c.dim;
    switch (c.eqn) {
    case DPB_EQN_LQR:
        e.k = (sqrt(c.discount * c.discount * c.q * c.q + 4.0 * c.p * c.q * (c.beta * c.beta)) - c.q * c.discount)
              / (c.beta * c.beta) / 2.0;                                  // equation.py:151
        e.cu = -c.beta * e.k / c.q;
        e.wconst = 2.0 * e.k * d;
        e.ZR = e.k * (c.R * c.R);
        break;
    case DPB_EQN_VDP:
        e.wconst = 2.0 * c.a * d;
        break;
    case DPB_EQN_EKN: {
        const double epsl = 1.0 / 2.0 / c.a2 / d;                         // equation.py:246
        if (c.ekn_sigma_fix) e.sig = sqrt(2.0 * epsl);
        e.C0 = 3.0 * (d + 1.0) * c.a3 / 2.0 / c.a2 / d;                   // equation.py:272
        e.gamma = c.discount;
        break;
    }
    default:
        e.k = (sqrt(5.0) - 1.0) / 2.0;                                    // equation.py:282
        e.wconst = 2.0 * e.k * d;
        e.ZR = e.k * (c.R * c.R);
        e.lv_num = e.k * e.k * ((c.beta + 2.0 * c.epsilon) * (c.beta + 2.0 * c.epsilon));
        e.lv_den = 2.0 * e.k * (c.epsilon * c.epsilon);
        e.lv_un = c.beta + 2.0 * c.epsilon;
        e.lv_ud = c.q / e.k;
        e.lv_ue = 2.0 * (c.epsilon * c.epsilon);
        e.lv_gk = c.discount * e.k;
    }
    e.delta_t = T / (double)N;                                            // equation.py:48,75
    e.sqrt_delta_t = sqrt(e.delta_t);                                     // equation.py:49
    e.hb = sigU * sqrt(3.0 * d * e.delta_t);                              // equation.py:80
    e.c3 = (3.0 * d) * (sigU * sigU);                                     // equation.py:85
    e.hmin = e.delta_t * 1e-4;                                            // equation.py:86
}

}  // namespace dpb
