// Kernel parameters of the batched force MPC (kf_mpc.cuh); separate so that kf_abi.cu can fill them without seeing the kernel.
#pragma once

#include <stdint.h>

namespace okf {

struct MpcParams {
    long long N;
    int max_legs;                               // 1..4: bound on the legs not in swing of any problem (sizes shared memory)
    const double *x, *body_ref, *p, *contact;  // [12][N], [NH*12][N], [12][N], [4][N]
    double *forces;                             // [NH*12][N]
    uint32_t *status;                           // [N] optional
    double dt, inv_mass, inv_inertia[3], gravity, mu, fz_max, w_state[12], w_force;
};

}  // namespace okf
