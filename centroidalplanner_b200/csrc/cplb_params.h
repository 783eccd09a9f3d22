// cplb_params.h -- the read-only parameter block handed BY VALUE to every kernel launch
// (__grid_constant__, lives in the constant bank: one broadcast read per warp, no HBM traffic,
// no staleness between a setter and the next launch on any stream).
//
// It holds what the reference keeps scattered over its IFOPT objects: CentroidalStatics::_m/_g/
// _wrench_manip (CentroidalStatics.cpp:12-15), FrictionCone::_F_thr (FrictionCone.cpp:14),
// EnvironmentClass::_mu (Environment.h:46), Ground::_ground_z, Superquadric::_C/_R/_P, and the
// refs/weights of MinimizeCentroidalVariables (MinimizeCentroidalVariables.cpp:11-26).
#ifndef CPLB_PARAMS_H
#define CPLB_PARAMS_H

#include <stdint.h>

#define CPLB_KMAX_CONTACTS 32

#define CPLB_WANT_G 1u
#define CPLB_WANT_J 2u
#define CPLB_WANT_COST 4u
#define CPLB_WANT_GRAD 8u
#define CPLB_JAC_COMPUTED_K 64u  // instance-major: like PACKED, also without the slots that are copies +-x[col] (JacMap<ENV, 2>)
#define CPLB_JAC_PACKED_K 32u   // instance-major: the Jacobian slice holds only the x-dependent slots (JacMap<ENV, true>)
#define CPLB_INPUTS_READY 16u  // x / per-instance arrays are not outputs of the preceding kernel: read them before griddepcontrol.wait

struct CplbParams {
    int32_t nc;
    int32_t env;
    int32_t n, m, nnz;
    int32_t reduction_order;            // Eigen fixed-size 3-term reductions: 0 = (v0+v1)+v2, 1 = v0+(v1+v2)
    int32_t perm[CPLB_KMAX_CONTACTS];  // sorted-name rank -> index in the caller's vector
    double mg[3];                       // _m * _g, one IEEE multiply per component, done on the host
    double mass;                        // _m (per-instance mode recomputes _m * _g on the device)
    double wrench[6];
    double mu;
    double ground_z;
    // Superquadric: raw parameters plus the constant-argument pow() results the reference
    // recomputes on every call (host glibc pow on the same arguments gives the same bits).
    double sqC[3], sqR[3], sqP[3];
    double sqPoverRP[3];  // P / pow(R, P)          Superquadric.cpp:54-56
    double sqRmP[3];      // pow(R, -P)             :98,109,...
    double sqRm2P[3];     // pow(R, -(P*2.0))       :82,86,107,...
    double sqR2P[3];      // pow(R,  P*2.0)         :95,96,149,150,203,204
    int32_t sqIntP[3];    // P as an integer when all three curvatures are integers in [2, 63], else 0
    int32_t sqBits;       // chain length: bit length of max(P) - 2
    int32_t sqWindow;     // |binary exponent of d| <= window keeps d^(2P) a normal number
    int32_t pad1;
    double F_thr[CPLB_KMAX_CONTACTS];
    double com_ref[3];
    double W_com;
    double p_ref[CPLB_KMAX_CONTACTS][3];
    double F_ref[CPLB_KMAX_CONTACTS][3];
    double W_p[CPLB_KMAX_CONTACTS];
    double W_F[CPLB_KMAX_CONTACTS];
};

// Optional per-instance parameter arrays (device memory; NULL = use the shared value of CplbParams), see
// cplb_instance_params in include/cpl_batched.h.  Same layout rule as x: element e of instance i is arr[i*len + e]
// (instance-major) or arr[e*ld + i] (component-major); len = 1, 6, nc, 3, 3*nc as noted.
struct CplbInstParams {
    const double* mass;      // 1
    const double* wrench;    // 6
    const double* mu;        // 1
    const double* F_thr;     // nc   (contact index = position in the caller's name vector)
    const double* ground_z;  // 1
    const double* com_ref;   // 3
    const double* W_com;     // 1
    const double* p_ref;     // 3*nc (x,y,z of contact 0, then contact 1, ...)
    const double* F_ref;     // 3*nc
    const double* W_p;       // nc
    const double* W_F;       // nc
};

// Instance-major kernel, per-instance parameter mode: where each staged parameter array's slice of a warp tile sits in the
// shared-memory parameter buffer (doubles from the buffer start; -1 = not staged: absent, or not needed by the requested
// outputs).  Array order = the member order of CplbInstParams.
#define CPLB_NUM_INST_ARRAYS 11
struct CplbParamTile {
    int off[CPLB_NUM_INST_ARRAYS];
    int len[CPLB_NUM_INST_ARRAYS];  // doubles per instance of each array
    int total;                      // doubles per buffer (one warp tile)
};

// Pointers of one evaluation (device memory), see cplb_eval_args in include/cpl_batched.h.
struct CplbIo {
    const double* x;
    double* g;
    double* jac;
    double* cost;
    double* grad;
    long long ld;  // COMPONENT_MAJOR row pitch (elements)
    long long N;
};

#endif
